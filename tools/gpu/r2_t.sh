mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
WL=c3 bash tools/gpu/ncu_skew.sh
