# ncu evidence for the default workload (BASELINE configs[1]): full capture of one step's kernels and
# the launch list.  Run under gpurun; results land in gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
python bench.py --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 15 -c 5 -o gpurun_out/prof_c2 -f \
    python bench.py --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/ncu_full_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/ncu_list_c2.log 2>&1
tail -1 gpurun_out/ncu_full_c2.log | cut -c1-200
