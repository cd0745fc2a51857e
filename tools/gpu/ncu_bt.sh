mkdir -p gpurun_out
python bench.py --workload c3 --no-cpu --steps 3 --warmup 3 > gpurun_out/plain_c3bt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_backtrace -s 3 -c 1 -o gpurun_out/prof_c3_bt -f python bench.py --workload c3 --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_c3bt.log 2>&1
tail -2 gpurun_out/ncu_c3bt.log
