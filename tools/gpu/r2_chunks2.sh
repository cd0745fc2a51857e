mkdir -p gpurun_out
run() { # corpus cc ns
  timeout 300 python bench.py --gpus 1 --workload c5 --corpus $1 --no-extra --no-e2e --steps 10 --corpus-chunk-cells $2 --corpus-streams $3 > gpurun_out/chunks2.json 2> gpurun_out/chunks2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/chunks2.json").read().strip().splitlines()[-1])
    print("corpus $1 chunk $2 streams $3: ms %.4f value %.4g chunks %s" % (d["ms_per_step"], d["value"], d["run_config"]["chunks_per_rank"]))
except Exception as e: print("corpus $1 chunk $2 streams $3 ERR", e, open("gpurun_out/chunks2.err").read()[-300:])
PY
}
run 4096 300000000 2
run 4096 400000000 2
run 8192 300000000 2
run 8192 700000000 2
run 32768 150000000 2
run 32768 300000000 2
run 32768 600000000 2
run 32768 1200000000 2
run 32768 300000000 3
