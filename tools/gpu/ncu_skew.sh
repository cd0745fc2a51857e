# ncu --set full capture of the skewed DP kernel: WL=c3 (one warp per SM: a clean single-warp picture) or c2
mkdir -p gpurun_out
WL=${WL:-c3}
python bench.py --workload $WL --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/plain_skew.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_dp_skew -s 3 -c 1 -o gpurun_out/prof_skew_$WL -f \
    python bench.py --workload $WL --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/ncu_skew_$WL.log 2>&1
tail -2 gpurun_out/ncu_skew_$WL.log | cut -c1-300
