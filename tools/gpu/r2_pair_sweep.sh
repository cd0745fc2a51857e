# pair layout: unroll depth x launch order (library variants built into hubertfa_b200/csrc/build/var)
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
cp hubertfa_b200/libhfa_align.so /tmp/libhfa_main.so
for u in ${UNROLLS:-2 4 8}; do
  cp hubertfa_b200/csrc/build/var/libhfa_align_u$u.so hubertfa_b200/libhfa_align.so
  for ord in new old; do
    if [ $ord = old ]; then export HFA_WARP_ORDER=old; else unset HFA_WARP_ORDER; fi
    for p in ${PAIRS:-2}; do
      HFA_PAIR=$p timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > $O/sweep_u${u}_${ord}_p$p.json 2> $O/sweep_u${u}_${ord}_p$p.err
    done
  done
done
unset HFA_WARP_ORDER
cp /tmp/libhfa_main.so hubertfa_b200/libhfa_align.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p/sweep_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], "dp %.4f"%r["stage_ms"]["dp"], "frac %.3f"%r["frac"], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
