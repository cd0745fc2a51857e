mkdir -p gpurun_out/r2h
timeout 300 python tools/prof_pipeline.py > gpurun_out/r2h/pipe.log 2>&1
cat gpurun_out/r2h/pipe.log
