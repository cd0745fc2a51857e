mkdir -p gpurun_out
for v in "" "HFA_DP_MAXK8=1" "HFA_DP_WARPS_PER_SM=13" "HFA_DP_MAXK8=1 HFA_DP_WARPS_PER_SM=13"; do
env $v timeout 300 python bench.py --workload c4 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print("[%s]"%sys.argv[1], "c4 ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"])
PY
done
