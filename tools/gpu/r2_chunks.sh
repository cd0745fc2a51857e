mkdir -p gpurun_out
for cc in 150000000 100000000 75000000 50000000 37500000; do
 for ns in 2 3 4; do
  timeout 200 python bench.py --gpus 1 --workload c5 --corpus 4096 --no-extra --no-e2e --steps 20 --corpus-chunk-cells $cc --corpus-streams $ns > gpurun_out/chunks_${cc}_$ns.json 2> gpurun_out/chunks_${cc}_$ns.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/chunks_${cc}_$ns.json").read().strip().splitlines()[-1])
    print("chunk $cc streams $ns: ms %.4f value %.4g chunks %s launches %d" % (d["ms_per_step"], d["value"], d["run_config"]["chunks_per_rank"], d["gpu_launches"]))
except Exception as e: print("chunk $cc streams $ns ERR", e)
PY
 done
done
