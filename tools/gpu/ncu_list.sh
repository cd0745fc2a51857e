mkdir -p gpurun_out
python bench.py --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/plain_list.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --no-cpu --no-extra --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
