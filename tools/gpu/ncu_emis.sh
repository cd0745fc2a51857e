mkdir -p gpurun_out
python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_c4e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_emission_stream -s 3 -c 1 -o gpurun_out/prof_c4_emis -f python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_c4e.log 2>&1
tail -1 gpurun_out/ncu_c4e.log | cut -c1-200
