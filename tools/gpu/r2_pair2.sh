# pair layout as the default for dictionary-style big batches: whole GPU suite, stage-count variants, ncu
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/gpu_tests.log
cp hubertfa_b200/libhfa_align.so /tmp/libhfa_main.so
for v in s3 s2; do
  cp hubertfa_b200/csrc/build/var/libhfa_align_$v.so hubertfa_b200/libhfa_align.so
  timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > $O/v_${v}_c4.json 2> $O/v_${v}_c4.err
  HFA_PAIR=0 timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 20 > $O/v_${v}_c4_p0.json 2> $O/v_${v}_c4_p0.err
done
cp /tmp/libhfa_main.so hubertfa_b200/libhfa_align.so
timeout 300 python bench.py --workload c4j --no-cpu --no-extra --steps 20 > $O/v_s3_c4j.json 2> $O/v_s3_c4j.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p/v_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["stage_ms"].items()}, "frac %.3f"%r["frac"], d["verified"]["paths_equal_to_oracle"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
PAIRS="1" bash tools/gpu/r2_pair_ncu.sh
