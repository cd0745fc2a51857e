mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
for a in 0 0.15 0.26 0.4 0.6; do
  HFA_HYBRID=$a timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > gpurun_out/bench_c4_h$a.json 2> gpurun_out/bench_c4_h$a.err
done
HFA_HYBRID=0 timeout 300 python bench.py --workload m1024 --no-cpu --no-extra --steps 50 > gpurun_out/bench_m1024_h0.json 2> gpurun_out/bench_m1024_h0.err
timeout 300 python bench.py --workload m1024 --no-cpu --no-extra --steps 50 > gpurun_out/bench_m1024_h0.26.json 2> gpurun_out/bench_m1024_h0.26.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4_h*.json"))+sorted(glob.glob("gpurun_out/bench_m1024_h*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["stage_ms"].items()}, "frac %.3f"%r["frac"], r["kernel"][:150], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e)
PY
