mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core_parity.py -x -q -k "not c3" > gpurun_out/t_band.log 2>&1; echo "rc=$?" >> gpurun_out/t_band.log
tail -3 gpurun_out/t_band.log
for v in desc asc; do
  if [ $v = asc ]; then cp tools/gpu/lib_asc.so hubertfa_b200/libhfa_align.so; fi
  for r in 1 2; do
  timeout 300 python bench.py --no-cpu --no-extra --steps 100 > gpurun_out/bench_c2_$v$r.json 2> gpurun_out/bench_c2_$v.err
  done
  for k in 2 4; do HFA_BIG_K=$k timeout 300 python bench.py --workload c3 --no-cpu --steps 20 > gpurun_out/bench_c3_$v$k.json 2> gpurun_out/bench_c3_$v.err; done
done
python - <<'PY'
import json
for f in ["c2_desc1","c2_desc2","c2_asc1","c2_asc2","c3_desc2","c3_asc2","c3_desc4","c3_asc4"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
