"""Host-side time breakdown of hubertfa_b200.pipeline.HostBatchAligner on the config-2 batch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hubertfa_b200 import synth, ops, _lib
from hubertfa_b200.pipeline import BufferPool, HostBatchAligner

dev = torch.device("cuda")
T, S, V, _ = bench.workload_shapes("c2", synth.SEED0)
ids = np.concatenate(synth.make_ids_batch(T, S, V, seed=synth.SEED0))
head = bench.make_head(T, V, 1).pin_memory()
pool = BufferPool(dev)
def tm(f, n=20):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for ch in (None, 1, 4, 8):
    al = HostBatchAligner(T, S, ids, V, 0.02, V + 2, n_chunks=ch, device=dev, pool=pool)
    print(f"chunks={ch}: run {tm(lambda: al.run(head)):.3f} ms")
    torch.cuda.synchronize()
    print("   ", [{k: round(v, 3) for k, v in c.items()} for c in al.run(head, profile=True)["profile"]])
# raw pieces
dh = torch.empty_like(head, device=dev)
print("H2D one copy", tm(lambda: dh.copy_(head, non_blocking=True)), "ms")
plan = ops.AlignPlan(T, S, ids, V, 0.02)
print("AlignPlan create", tm(lambda: ops.AlignPlan(T, S, ids, V, 0.02)), "ms")
ws = plan.new_workspace(dev); res = plan.new_result(dev)
print("upload", tm(lambda: plan.upload(ws)), "ms")
row0 = plan.frame_off[:-1]
def si(base):
    plan.set_inputs(ws, base + (row0 * (V + 2) + 2) * 4, np.full(plan.n_utt, V + 2), np.ones(plan.n_utt), base + row0 * (V + 2) * 4, np.full(plan.n_utt, V + 2))
print("set_inputs", tm(lambda: si(dh.data_ptr())), "ms")
si(dh.data_ptr())
print("emission from device", tm(lambda: ops.emission(ws, plan.handle, 0)), "ms")
si(head.data_ptr())
print("emission zero-copy from pinned host", tm(lambda: ops.emission(ws, plan.handle, 0)), "ms")
print("align_batch zero-copy", tm(lambda: ops.align_batch(ws, plan.handle, 0, res, None)), "ms")
si(dh.data_ptr())
print("align_batch device", tm(lambda: ops.align_batch(ws, plan.handle, 0, res, None)), "ms")
