mkdir -p gpurun_out
for n in 0 6 8 10 12 16 24; do
HFA_DP_WARPS_PER_SM=$n timeout 300 python bench.py --workload c4 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - $n <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("warps/SM cap", sys.argv[1], "c4 ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e %.3f"%d["e2e"]["ms_per_step"])
PY
done
