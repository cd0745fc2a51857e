mkdir -p gpurun_out/r2
TOOL=${TOOL:-memcheck}
python tools/gpu/sanitize_case.py > gpurun_out/r2/sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/gpu/sanitize_case.py > gpurun_out/r2/sanitize_$TOOL.log 2>&1
echo "sanitizer rc=$?"; tail -15 gpurun_out/r2/sanitize_$TOOL.log
