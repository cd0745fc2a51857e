# round-2 final measurements on one B200 (plain runs first, then the ncu passes of the same commands)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_ref.err
timeout 300 python bench.py --workload c1 --no-extra --steps 300 --cpu-seconds 3 > $O/bench_c1.json 2> $O/bench_c1.err
timeout 300 python bench.py --workload c3 --no-extra --steps 30 --cpu-seconds 5 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --workload c4 --no-extra --no-cpu --steps 20 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --workload c4j --no-extra --no-cpu --steps 20 > $O/bench_c4j.json 2> $O/bench_c4j.err
timeout 300 python tools/time_decode.py > $O/time_decode.log 2>&1; tail -3 $O/time_decode.log
# ncu: launch list, then full captures of one step's kernels (eager launches so that every kernel is a launch)
python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_c2.csv \
    python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/ncu_list_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 15 -c 5 -o $O/prof_c2 -f \
    python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/ncu_full_c2.log 2>&1
tail -1 $O/ncu_full_c2.log | cut -c1-160
python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 12 -c 4 -o $O/prof_c4 -f \
    python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/ncu_full_c4.log 2>&1
tail -1 $O/ncu_full_c4.log | cut -c1-160
python - <<'PY'
import json
O="gpurun_out/r2"
for f in ["c2","reference_arm","c1","c3","c4","c4j"]:
    try:
        d=json.loads(open(f"{O}/bench_{f}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g (%.3f ms)"%(d["e2e"]["value"], d["e2e"].get("ms_per_step",0)), r.get("stage_ms"), "frac", r.get("frac"), "kept", (r.get("with_kept_dp") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("verified") or {}).get("paths_equal_to_oracle"))
        if "extra" in d:
            print("   c4", d["extra"]["c4"]["ms_per_step"], d["extra"]["c4"]["roofline"]["stage_ms"], d["extra"]["c4"]["roofline"]["frac"])
            print("   corpus", {k:d["extra"]["corpus"][k] for k in ("value","ms_per_step","chunks","all_status_ok")})
    except Exception as e: print(f, "ERR", e)
PY
