# round-2 regression check: GPU test suite (core first), then short bench lines per config
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core_parity.py -m gpu -x -q > gpurun_out/t_core.log 2>&1; echo "rc=$?" >> gpurun_out/t_core.log
tail -5 gpurun_out/t_core.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_core_parity.py -s > gpurun_out/t_rest.log 2>&1; echo "rc=$?" >> gpurun_out/t_rest.log
tail -8 gpurun_out/t_rest.log
for wl in c2 c1 c3; do
  st=200; [ $wl = c3 ] && st=20
  timeout 300 python bench.py --workload $wl --no-cpu --no-extra --steps $st > gpurun_out/bench_${wl}_a.json 2> gpurun_out/bench_${wl}_a.err
  HFA_LAT_KERNEL=band timeout 300 python bench.py --workload $wl --no-cpu --no-extra --steps $st > gpurun_out/bench_${wl}_band.json 2> gpurun_out/bench_${wl}_band.err
  HFA_SKEW_D=3 timeout 300 python bench.py --workload $wl --no-cpu --no-extra --steps $st > gpurun_out/bench_${wl}_d3.json 2> gpurun_out/bench_${wl}_d3.err
done
python - <<'PY'
import json
for f in ["c2_a","c2_band","c2_d3","c1_a","c1_band","c1_d3","c3_a","c3_band","c3_d3"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
