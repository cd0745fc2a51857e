# routing crossover: the strip kernel (HFA_LATENCY_MODE=2) vs one warp per utterance (=0) on batches between configs[1] and configs[3]
mkdir -p gpurun_out
for wl in c2 m512 m1024 m2048; do
  for mode in 0 2; do
    HFA_LATENCY_MODE=$mode timeout 300 python bench.py --workload $wl --no-cpu --no-extra --steps 50 > gpurun_out/cross_${wl}_$mode.json 2> gpurun_out/cross_${wl}_$mode.err
  done
done
python - <<'PY'
import json
rows=[]
for wl in ["c2","m512","m1024","m2048"]:
    for mode in ["0","2"]:
        try:
            d=json.loads(open(f"gpurun_out/cross_{wl}_{mode}.json").read().strip().splitlines()[-1])
            r=d["roofline"]
            rows.append(dict(workload=wl, routing=("one warp per utterance" if mode=="0" else "strips (skewed kernel)"), utterances=d["config"]["utterances"], ms_per_step=round(d["ms_per_step"],4), stage_ms={k:round(v,4) for k,v in r["stage_ms"].items()}, kernel=r["kernel"][:90]))
            print(rows[-1])
        except Exception as e: print(wl, mode, "ERR", e)
json.dump(rows, open("gpurun_out/r2_routing_crossover.json","w"), indent=1)
PY
