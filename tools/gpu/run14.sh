mkdir -p gpurun_out
cp hubertfa_b200/libhfa_align.so /tmp/lib_v0.so
for v in 0 1 2 3; do
  cp /tmp/lib_v0.so hubertfa_b200/libhfa_align.so
  if [ $v != 0 ]; then cp tools/gpu/lib_v$v.so hubertfa_b200/libhfa_align.so; fi
  timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  timeout 300 python bench.py --no-extra --no-cpu --steps 100 > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
  python - $v <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
e=json.loads(open("gpurun_out/bench_y.json").read().strip().splitlines()[-1])
print("variant", sys.argv[1], "c3 dp ms %.3f"%d["roofline"]["stage_ms"]["dp"], " c2 dp ms %.4f"%e["roofline"]["stage_ms"]["dp"])
PY
done
