mkdir -p gpurun_out
K=${K:-hfa_dp_warp_any}
python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_c4_$K -f \
    python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_c4_$K.log 2>&1
tail -2 gpurun_out/ncu_c4_$K.log | cut -c1-200
