set -x
O=gpurun_out/r2g
mkdir -p $O
for w in c4 c2; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $O/list_$w.csv python bench.py --workload $w --no-cpu --no-extra --no-e2e --no-graph --steps 2 --warmup 1 > $O/list_$w.log 2>&1
done
python - <<'PY'
import csv
for w in ["c4","c2"]:
    rows=[r for r in csv.reader(open(f"gpurun_out/r2g/list_{w}.csv")) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
    d={}
    for r in rows[1:]:
        d.setdefault((int(r[ii]),r[ki][:60]),{})[r[mi]]=r[vi]
    for k in sorted(d)[-14:]:
        print(w,k,d[k])
PY
