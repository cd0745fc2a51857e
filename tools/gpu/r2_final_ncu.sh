# round-2 final ncu passes (after tools/gpu/r2_final.sh): launch list, full captures of one step's kernels
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
# ncu: launch list, then full captures of one step's kernels (eager launches so that every kernel is a launch)
python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_c2.csv \
    python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/ncu_list_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 15 -c 5 -o $O/prof_c2 -f \
    python bench.py --no-cpu --no-extra --no-graph --steps 3 --warmup 3 > $O/ncu_full_c2.log 2>&1
tail -1 $O/ncu_full_c2.log | cut -c1-160
python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 12 -c 4 -o $O/prof_c4 -f \
    python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/ncu_full_c4.log 2>&1
tail -1 $O/ncu_full_c4.log | cut -c1-160
