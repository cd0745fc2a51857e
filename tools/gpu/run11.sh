mkdir -p gpurun_out
HFA_BAND_MAX=0 HFA_BAND_TOP=0.3 timeout 600 python -m pytest tests/test_gpu_core_parity.py -x -q -k "auto and not c3" > gpurun_out/t_top.log 2>&1; echo "rc=$?" >> gpurun_out/t_top.log
tail -3 gpurun_out/t_top.log
for f in 0 0.02 0.05 0.1 0.2 0.4; do
HFA_BAND_TOP=$f timeout 300 python bench.py --workload c4 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - $f <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("band top", sys.argv[1], "c4 ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"])
PY
done
