#!/usr/bin/env python
"""Times the UNMODIFIED reference decoder (Python + numba, imported from /root/reference) on a sample of
bench.py's default workload and writes profiles/reference_python_cpu.json.  Runs where the reference tree
exists (the build container, no GPU); bench.py prints the file's content as `cpu_baseline.reference_python`
beside the C port it times on the GPU box's host cores."""
import json
import os
import platform
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_head, workload_shapes          # noqa: E402
from hubertfa_b200 import synth                        # noqa: E402
from oracle import c_oracle as oc                      # noqa: E402
from oracle.reference_import import load_reference_decoder   # noqa: E402


def main():
    Ref = load_reference_decoder()
    if Ref is None:
        raise SystemExit("the reference tree (/root/reference) or numba is not available here")
    torch.set_num_threads(1)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T, S, V, desc = workload_shapes("c2", synth.SEED0)
    ids_list = synth.make_ids_batch(T, S, V, seed=synth.SEED0)
    sel = np.unique(np.linspace(0, len(T) - 1, n).astype(np.int64))
    head = make_head(T, V, synth.SEED0)
    row_off = np.concatenate([[0], np.cumsum(T.astype(np.int64))])
    vocab = synth.make_vocab(V)
    names = np.array(["SP"] + [f"p{i}" for i in range(1, V)])
    dec = Ref(vocab["vocab"] and vocab, synth.MELSPEC_50FPS)
    items = []
    for b in sel:
        h = head[row_off[b]:row_off[b + 1]]
        items.append((h[None, :, 2:], h[None, :, 0], torch.zeros(1, int(T[b]), V), list(names[ids_list[b]])))
    dec.decode(*items[0][:3], None, items[0][3])       # numba JIT (~3 s), excluded
    t0 = time.perf_counter()
    for fr, ed, ctc, seq in items:
        dec.decode(fr, ed, ctc, None, seq)
    dt = time.perf_counter() - t0
    cells = int((T[sel].astype(np.int64) * S[sel]).sum())
    # the C port on the same sample, same box, one thread -- the ratio is what carries over to other hosts
    sub = np.ascontiguousarray(head.numpy()[np.concatenate([np.arange(row_off[b], row_off[b + 1]) for b in sel])])
    ids_cat = np.concatenate([ids_list[b] for b in sel])
    oc.align_batch(T[sel], S[sel], V, sub[:, 2:], sub[:, 0], ids_cat, synth.FRAME_SECONDS, 1)
    t0 = time.perf_counter()
    oc.align_batch(T[sel], S[sel], V, sub[:, 2:], sub[:, 0], ids_cat, synth.FRAME_SECONDS, 1)
    dt_c = time.perf_counter() - t0
    out = {"what": "unmodified tools/alignment_decoder.py AlignmentDecoder.decode (Python + numba), one thread",
           "sample": f"{len(sel)} utterances evenly spaced through bench.py's c2 batch ({cells} cells), JIT warm-up excluded",
           "value": cells / dt, "unit": "cells/s", "seconds": dt, "cores": 1,
           "c_port_same_sample_same_box": {"value": cells / dt_c, "unit": "cells/s", "cores": 1},
           "port_over_reference": (cells / dt_c) / (cells / dt),
           "where": f"build container ({platform.processor() or platform.machine()}, {os.cpu_count()} vCPU; no GPU): "
                    "the reference tree does not travel to the GPU box",
           "made_by": "tools/time_reference_python.py"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "reference_python_cpu.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
