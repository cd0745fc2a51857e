mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke.log
timeout 300 python tools/time_decode.py > gpurun_out/r2/time_decode.log 2>&1; head -1 gpurun_out/r2/time_decode.log
