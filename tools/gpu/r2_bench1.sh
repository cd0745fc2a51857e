mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q > gpurun_out/t_api.log 2>&1; echo "rc=$?" >> gpurun_out/t_api.log
tail -15 gpurun_out/t_api.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 300 python bench.py --workload c1 --no-extra --no-cpu --steps 300 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 300 python bench.py --workload c1 --no-extra --no-cpu --no-graph --steps 300 > gpurun_out/bench_c1_eager.json 2> gpurun_out/bench_c1_eager.err
timeout 300 python bench.py --workload c3 --no-extra --no-cpu --steps 30 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python - <<'PY'
import json
for f in ["c2","reference_arm","c1","c1_eager","c3"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d.get("roofline",{}).get("stage_ms"), "frac", d.get("roofline",{}).get("frac"), "verified", d.get("verified"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if "extra" in d:
            print("   c4", d["extra"]["c4"]["ms_per_step"], d["extra"]["c4"]["roofline"]["stage_ms"], d["extra"]["c4"]["roofline"]["frac"])
            print("   corpus", d["extra"]["corpus"])
    except Exception as e: print(f, "ERR", e)
PY
