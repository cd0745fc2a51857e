# compacted emission rows (one column per distinct id) for the pair layout: whole GPU suite, config 4 with / without
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/gpu_tests.log
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > $O/c4.json 2> $O/c4.err
HFA_COMPACT=0 timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > $O/c4_plainrows.json 2> $O/c4_plainrows.err
timeout 300 python bench.py --workload c4j --no-cpu --no-extra --steps 20 > $O/c4j.json 2> $O/c4j.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c/c4*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["stage_ms"].items()}, "frac %.3f"%r["frac"], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
