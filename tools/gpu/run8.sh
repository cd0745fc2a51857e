mkdir -p gpurun_out
cp hubertfa_b200/libhfa_align.so /tmp/lib_keep.so
for v in keep nostore keep nostore; do
  cp /tmp/lib_keep.so hubertfa_b200/libhfa_align.so
  if [ $v = nostore ]; then cp tools/gpu/lib_nostore.so hubertfa_b200/libhfa_align.so; fi
  HFA_KEEP_DP=0 HFA_BIG_K=2 timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - $v <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print(sys.argv[1], "KEEP_DP=0 c3 dp ms", d["roofline"]["stage_ms"]["dp"])
PY
done
