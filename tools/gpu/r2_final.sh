# round-2 final measurements on one B200 (plain runs first, then the ncu passes of the same commands)
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_ref.err
timeout 300 python bench.py --workload c1 --no-extra --steps 300 --cpu-seconds 3 > $O/bench_c1.json 2> $O/bench_c1.err
timeout 300 python bench.py --workload c3 --no-extra --steps 30 --cpu-seconds 5 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --workload c4 --no-extra --no-cpu --steps 20 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --workload c4j --no-extra --no-cpu --steps 20 > $O/bench_c4j.json 2> $O/bench_c4j.err
timeout 300 python tools/time_decode.py > $O/time_decode.log 2>&1; tail -3 $O/time_decode.log
python - <<'PY'
import json
O="gpurun_out/r2z"
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
