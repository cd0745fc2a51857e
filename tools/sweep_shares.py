"""e2e sweep over the chunk shares of hubertfa_b200.pipeline.HostBatchAligner (config 2)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hubertfa_b200 import synth
from hubertfa_b200.pipeline import BufferPool, HostBatchAligner

dev = torch.device("cuda")
T, S, V, _ = bench.workload_shapes("c2", synth.SEED0)
ids = np.concatenate(synth.make_ids_batch(T, S, V, seed=synth.SEED0))
heads = [bench.make_head(T, V, i).pin_memory() for i in range(4)]
pool = BufferPool(dev)
def tm(f, n=100):
    for i in range(5): f(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
cands = {
    "default": None,
    "4x25": (0.25, 0.25, 0.25, 0.25),
    "6 tail4": (0.28, 0.26, 0.22, 0.13, 0.07, 0.04),
    "6 tail3": (0.30, 0.27, 0.21, 0.12, 0.07, 0.03),
    "5 tail5": (0.32, 0.28, 0.22, 0.13, 0.05),
    "4 tail6": (0.36, 0.32, 0.26, 0.06),
    "3 tail8": (0.50, 0.42, 0.08),
}
for name, sh in cands.items():
    def run(i, sh=sh):
        al = HostBatchAligner(T, S, ids, V, 0.02, V + 2, device=dev, pool=pool, shares=sh)
        return al.run(heads[i % 4])
    print(f"{name:10s} {tm(run):.3f} ms")
