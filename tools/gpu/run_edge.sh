O=gpurun_out/r2i
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_decode_parity.py -x -q -k "emission or golden" > $O/t.log 2>&1; tail -2 $O/t.log
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --no-e2e --steps 50 > $O/c4.json 2> $O/c4.err
timeout 300 python bench.py --workload c4j --no-cpu --no-extra --no-e2e --steps 50 > $O/c4j.json 2> $O/c4j.err
timeout 300 python bench.py --no-cpu --no-extra --no-e2e > $O/c2.json 2> $O/c2.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:edge -c 3 --csv --log-file $O/list_c4.csv python bench.py --workload c4 --no-cpu --no-extra --no-e2e --no-graph --steps 2 --warmup 1 > $O/list_c4.log 2>&1
python - <<'PY'
import json,csv
for f in ["c4","c4j","c2"]:
    d=json.loads(open(f"gpurun_out/r2i/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"])
rows=[r for r in csv.reader(open("gpurun_out/r2i/list_c4.csv")) if len(r)>10]
for r in rows[1:]: print(r[4][:40], r[-3], r[-1])
PY
