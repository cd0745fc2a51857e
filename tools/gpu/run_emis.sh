O=gpurun_out/r2e
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/t.log 2>&1; tail -2 $O/t.log
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --no-e2e --steps 50 > $O/c4.json 2> $O/c4.err
timeout 300 python bench.py --workload c4j --no-cpu --no-extra --no-e2e --steps 50 > $O/c4j.json 2> $O/c4j.err
timeout 300 python bench.py --no-cpu --no-extra --no-e2e > $O/c2.json 2> $O/c2.err
timeout 300 python bench.py --workload c1 --no-cpu --no-extra --no-e2e --steps 300 > $O/c1.json 2> $O/c1.err
python - <<'PY'
import json
for f in ["c4","c4j","c2","c1"]:
    d=json.loads(open(f"gpurun_out/r2e/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"], d["verified"]["paths_equal_to_oracle"])
PY
