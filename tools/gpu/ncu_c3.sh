mkdir -p gpurun_out
export HFA_BIG_KERNEL=band HFA_BIG_K=2
python bench.py --workload c3 --no-cpu --steps 3 --warmup 3 > gpurun_out/plain_c3b2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_dp_band -s 3 -c 1 -o gpurun_out/prof_c3_band2 -f python bench.py --workload c3 --no-cpu --steps 3 --warmup 3 > gpurun_out/ncu_c3b2.log 2>&1
tail -3 gpurun_out/ncu_c3b2.log
