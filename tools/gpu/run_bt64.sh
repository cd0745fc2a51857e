O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; tail -2 $O/t_all.log
timeout 300 python bench.py --workload c4 --no-cpu --no-extra --no-e2e --steps 50 > $O/c4b.json 2> $O/c4b.err
timeout 300 python bench.py --workload c3 --no-cpu --no-extra --no-e2e --steps 30 > $O/c3b.json 2> $O/c3b.err
timeout 300 python bench.py --workload c1 --no-cpu --no-extra --no-e2e --steps 300 > $O/c1b.json 2> $O/c1b.err
timeout 300 python bench.py --no-cpu --no-extra --no-e2e > $O/c2b.json 2> $O/c2b.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:backtrace -c 2 --csv --log-file $O/list_c4b.csv python bench.py --workload c4 --no-cpu --no-extra --no-e2e --no-graph --steps 2 --warmup 1 > $O/list_c4b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:backtrace -c 2 --csv --log-file $O/list_c2b.csv python bench.py --no-cpu --no-extra --no-e2e --no-graph --steps 2 --warmup 1 > $O/list_c2b.log 2>&1
python - <<'PY'
import json,csv
for f in ["c4b","c3b","c1b","c2b"]:
    d=json.loads(open(f"gpurun_out/r2i/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step %.4f"%d["ms_per_step"], d["roofline"]["stage_ms"])
for f in ["list_c4b","list_c2b"]:
    rows=[r for r in csv.reader(open(f"gpurun_out/r2i/{f}.csv")) if len(r)>10]
    for r in rows[1:]: print(r[4][:50], r[-3], r[-1])
PY
