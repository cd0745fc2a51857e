O=gpurun_out/r2e
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_decode_parity.py tests/test_gpu_full_size.py -x -q > $O/t.log 2>&1; tail -2 $O/t.log
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --no-e2e --steps 50 > $O/c4.json 2> $O/c4.err
timeout 300 python bench.py --workload c4j --no-cpu --no-extra --no-e2e --steps 50 > $O/c4j.json 2> $O/c4j.err
timeout 300 python bench.py --no-cpu --no-extra --no-e2e > $O/c2.json 2> $O/c2.err
cp hubertfa_b200/csrc/build/var/libhfa_align_fx.so hubertfa_b200/libhfa_align.so
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --no-e2e --steps 50 > $O/c4_fx.json 2> $O/c4_fx.err
timeout 300 python bench.py --no-cpu --no-extra --no-e2e > $O/c2_fx.json 2> $O/c2_fx.err
python - <<'PY'
import json
for f in ["c4","c4j","c2","c4_fx","c2_fx"]:
    d=json.loads(open(f"gpurun_out/r2e/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], d["verified"]["paths_equal_to_oracle"])
PY
