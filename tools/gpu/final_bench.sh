mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err
python bench.py --workload c1 --steps 300 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --workload c3 --steps 30 --cpu-seconds 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c4 --steps 20 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
for f in ["c2","reference_arm","c1","c3","c4"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d.get("roofline",{}).get("stage_ms"), "frac", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if "extra" in d: print("   c4", d["extra"]["c4"]["ms_per_step"], d["extra"]["c4"]["roofline"]["stage_ms"], d["extra"]["c4"]["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
