mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "rc=$?"
tail -5 gpurun_out/bench_c5_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
for f in ["c5_${N}gpu","ref_${N}gpu"]:
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, json.dumps({k:d[k] for k in d if k not in ("config",)}, indent=0)[:3000])
        print(d["run_config"].get("chunks_per_rank"), d["run_config"].get("cells_per_rank"))
    except Exception as e: print(f, "ERR", e)
PY
