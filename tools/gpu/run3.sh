mkdir -p gpurun_out
./tools/ubench_warpid | tee gpurun_out/ubench_warpid.txt
timeout 900 python -m pytest tests/test_gpu_core_parity.py -x -q > gpurun_out/t_band.log 2>&1; echo "rc=$?" >> gpurun_out/t_band.log
tail -3 gpurun_out/t_band.log
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/bench_c2_p.json 2> gpurun_out/bench_c2_p.err
HFA_DP_FORM=cu timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/bench_c2_cu.json 2> gpurun_out/bench_c2_cu.err
HFA_BIG_K=2 timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_c3_band2.json 2> gpurun_out/bench_c3_band2.err
python - <<'PY'
import json
for f in ["c2_p","c2_cu","c3_band2"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
        if "extra" in d: print("   c4", "ms/step %.4f"%d["extra"]["c4"]["ms_per_step"], d["extra"]["c4"]["roofline"]["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
