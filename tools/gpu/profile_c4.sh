# ncu evidence for the machine-filling workload (BASELINE configs[3], 4096 utterances)
mkdir -p gpurun_out
python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hfa_ -s 12 -c 4 -o gpurun_out/prof_c4 -f \
    python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_full_c4.log 2>&1
tail -1 gpurun_out/ncu_full_c4.log | cut -c1-200
