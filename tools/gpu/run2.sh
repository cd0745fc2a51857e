mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core_parity.py -x -q > gpurun_out/t_band.log 2>&1; echo "rc=$?" >> gpurun_out/t_band.log
tail -5 gpurun_out/t_band.log
timeout 300 python bench.py --no-extra --no-cpu --steps 100 > gpurun_out/bench_c2_band.json 2> gpurun_out/bench_c2_band.err
for k in 2 4; do HFA_BIG_KERNEL=band HFA_BIG_K=$k timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_c3_band$k.json 2> gpurun_out/bench_c3_band$k.err; done
python - <<'PY'
import json
for f in ["c2_band","c3_band2","c3_band4"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
