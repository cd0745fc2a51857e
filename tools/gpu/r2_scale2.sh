# strong scaling of the corpus arm at N = 2 (the N = 1 reference is measured inside the run)
mkdir -p gpurun_out
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "N=$N rc=$?"
tail -c 600 gpurun_out/bench_c5_${N}gpu.json
