# strong scaling of the corpus arm on one 8-GPU box: N = 8, 4, 2 (the N = 1 reference is measured inside every run)
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "N=$N rc=$?"
done
python - <<'PY'
import json
for N in (8,4):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_c5_{N}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(N, "value %.4g"%d["value"], "ms %.3f"%d["ms_per_step"], "n1", d.get("n1_same_workload"), "e2e %.4g"%d["e2e"]["value"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], d["e2e"]["h2d_ceiling"], "frac %.3f"%d["e2e"]["frac_of_h2d_ceiling"], d["run_config"]["chunks_per_rank"], d["verified"]["paths_equal_to_oracle"], d["extra"]["c2_replicas"]["value"], d["extra"]["c2_replicas"]["ms_per_step"], d["clocks"])
    except Exception as e: print(N, "ERR", e)
PY
