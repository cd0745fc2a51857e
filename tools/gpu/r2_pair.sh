# SP-aware pair layout of the warp kernel: parity first, then config 4 with and without it
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_core_parity.py -m gpu -x -q -k "pair or warp" > $O/t_pair.log 2>&1; echo "pair tests rc=$?"; tail -5 $O/t_pair.log
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --workload ${WL:-c4} --no-cpu --no-extra --steps 20 > $O/bench_$n.json 2> $O/bench_$n.err
}
run c4_p0 HFA_PAIR=0
run c4_p0_oldorder HFA_PAIR=0 HFA_WARP_ORDER=old
run c4_p1 HFA_PAIR=1
run c4_p1_oldorder HFA_PAIR=1 HFA_WARP_ORDER=old
run c4_p2 HFA_PAIR=2
WL=c4j run c4j_p1 HFA_PAIR=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["stage_ms"].items()}, "frac %.3f"%r["frac"], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
PAIRS="1" bash tools/gpu/r2_pair_ncu.sh
