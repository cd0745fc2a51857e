# the per-rank share of the corpus at N = 8 (12 500 utterances, 9.3e8 cells) on ONE GPU: chunk size x streams
O=gpurun_out/r2s
mkdir -p $O
for cc in 600000000 320000000 240000000 160000000; do
for st in 1 2 3; do
timeout 300 python bench.py --workload c5 --corpus 12500 --corpus-chunk-cells $cc --corpus-streams $st --no-e2e --no-cpu --no-extra --steps 20 --warmup 3 > $O/s_${cc}_$st.json 2> $O/s_${cc}_$st.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s/s_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.3f"%d["ms_per_step"], "value %.4g"%d["value"], d["run_config"]["chunks_per_rank"])
    except Exception as e: print(f, "ERR", e)
PY
