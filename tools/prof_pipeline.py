"""Host-side time breakdown of hubertfa_b200.pipeline.HostBatchAligner on the config-2 batch: per piece, when
its DMA landed, when its results were back, when the host had issued it; and the run time for a few share
vectors of the upload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hubertfa_b200 import synth
from hubertfa_b200.pipeline import BufferPool, HostBatchAligner

dev = torch.device("cuda")
T, S, V, _ = bench.workload_shapes("c2", synth.SEED0)
ids = np.concatenate(synth.make_ids_batch(T, S, V, seed=synth.SEED0))
head = bench.make_head(T, V, 1).pin_memory()
pool = BufferPool(dev)


def tm(f, n=30):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


dh = torch.empty_like(head, device=dev)
print("H2D one copy  med/min ms", tm(lambda: dh.copy_(head, non_blocking=True)))
for shares in (None, (0.30, 0.28, 0.22, 0.12, 0.08), (0.34, 0.30, 0.22, 0.10, 0.04), (0.4, 0.3, 0.2, 0.07, 0.03),
               (0.3, 0.25, 0.2, 0.13, 0.08, 0.04), (0.5, 0.3, 0.15, 0.05), (0.6, 0.3, 0.1), (1.0,)):
    al = HostBatchAligner(T, S, ids, V, 0.02, V + 2, device=dev, pool=pool, shares=shares)
    med, mn = tm(lambda: al.run(head))
    print(f"shares={shares}: run med {med:.3f} min {mn:.3f} ms")
    prof = al.run(head, profile=True)["profile"]
    print("    ", [{k: round(v, 3) for k, v in c.items()} for c in prof])
