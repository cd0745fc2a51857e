mkdir -p gpurun_out
for a in 0.2 0.3 0.45; do
  HFA_HYBRID_KIND=cta HFA_HYBRID=$a timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > gpurun_out/bench_c4_cta$a.json 2> gpurun_out/bench_c4_cta$a.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4_cta*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["stage_ms"].items()}, "frac %.3f"%r["frac"], r["kernel"][:150], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-300:])
PY
