# GPU regression check used during development: the whole GPU test suite, then short bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 300 python bench.py --no-cpu --no-extra --steps 200 > gpurun_out/bench_c2_a.json 2> gpurun_out/bench_c2_a.err
timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_c3_a.json 2> gpurun_out/bench_c3_a.err
timeout 300 python bench.py --workload c1 --no-cpu --steps 200 > gpurun_out/bench_c1_a.json 2> gpurun_out/bench_c1_a.err
python - <<'PY'
import json
for f in ["c2_a","c3_a","c1_a"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
