#!/usr/bin/env python
"""BASELINE configs[4]: corpus-scale alignment -- N synthetic utterances (default 100 000, shapes as
configs[1]) sharded by utterance across the GPUs of one box, streamed through bounded workspaces.

    python tools/corpus_c5.py [--utterances 100000] [--chunk-cells 3e8]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/corpus_c5.py --gpus 8

STRONG scaling: the corpus is fixed, every rank takes the shard `sharding.shard_by_cost` gives it
(longest-processing-time-first over sum T*S), cuts it into chunks of at most --chunk-cells DP cells
(`sharding.chunk_by_bytes`) and aligns chunk after chunk with device-resident logits (generated on the
device, chunk by chunk: 100 000 utterances are 27 GB of f32 logits); results go to pinned host memory.
Prints one JSON line (rank 0): whole-job cells/s = corpus cells / max-over-ranks time, no collective
on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--utterances", type=int, default=100000)
    ap.add_argument("--chunk-cells", type=float, default=3e8)
    ap.add_argument("--vocab", type=int, default=63)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from hubertfa_b200 import _lib, ops, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    V = args.vocab
    T, S = synth.sample_shapes(args.utterances, seed=synth.SEED0, min_s=5, max_s=30, s_lo=20, s_hi=150)
    mine = sharding.shard_by_cost(T, S, world)[rank]
    Tm, Sm = T[mine], S[mine]
    # length-bucketed collation inside the shard: longest first, then consecutive chunks
    order = np.argsort(-Tm.astype(np.int64), kind="stable")
    Tm, Sm = np.ascontiguousarray(Tm[order]), np.ascontiguousarray(Sm[order])
    chunks = sharding.chunk_by_bytes(Tm, Sm, int(args.chunk_cells))
    ids_all = synth.make_ids_batch(Tm, Sm, V, seed=synth.SEED0 + rank)
    plans = []
    for c in chunks:
        plans.append(ops.AlignPlan(Tm[c], Sm[c], np.concatenate([ids_all[i] for i in c]), V, synth.FRAME_SECONDS))
    ws_bytes = max(p.workspace_bytes for p in plans)
    res_bytes = max(p.result_bytes for p in plans)
    rows = max(int(Tm[c].sum()) for c in chunks)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    res = torch.empty(res_bytes, dtype=torch.uint8, device=dev)
    head = torch.empty(rows, V + 2, dtype=torch.float32, device=dev)
    host = [torch.empty(p.result_bytes, dtype=torch.uint8, pin_memory=True) for p in plans]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def run_chunk(k, timed):
        p, c = plans[k], chunks[k]
        n_rows = int(Tm[c].sum())
        h = head[:n_rows]
        h.normal_(generator=g)                         # synthetic logits of this chunk, on the device
        h[:, 2:] *= 3.0
        h[:, 0] *= 2.0
        row0 = p.frame_off[:-1]
        st = h.stride(0)
        base = h.data_ptr()
        if timed is not None:
            timed[0].record()
        p.upload(ws)
        p.set_inputs(ws, base + (row0 * st + 2) * 4, np.full(p.n_utt, st), np.ones(p.n_utt), base + row0 * st * 4,
                     np.full(p.n_utt, st))
        ops.align_batch(ws, p.handle, _lib.DTYPE_F32, res[:p.result_bytes], None)
        host[k].copy_(res[:p.result_bytes], non_blocking=True)
        if timed is not None:
            timed[1].record()

    run_chunk(0, None)                                 # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plans]
    for k in range(len(plans)):
        run_chunk(k, evs[k])
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)        # alignment time only (logit synthesis excluded)
    bad = sum(int((p.views(hb.numpy())["status"] != 0).sum()) for p, hb in zip(plans, host))
    cells = float((Tm.astype(np.int64) * Sm).sum())
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([cells, float(bad)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        cells, bad = float(t[0].item()), int(t[1].item())
    if rank == 0:
        frames = float(T.astype(np.int64).sum())
        print(json.dumps({
            "metric": "dp_cells_per_s", "value": cells / (ms * 1e-3), "unit": "cells/s", "n_gpus": world,
            "ms_total": ms, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "audio_hours_per_s": frames * synth.FRAME_SECONDS / 3600 / (ms * 1e-3),
            "config": {"workload": f"c5: configs[4]: {args.utterances} utterances, 5-30 s, 20-150 phonemes, V={V}, "
                                   "sharded by utterance (LPT over sum T*S), no collective",
                       "chunks_on_rank0": len(plans), "chunk_cells": args.chunk_cells, "cells": cells,
                       "unaligned_utterances": bad}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
