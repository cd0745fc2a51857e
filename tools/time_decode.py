"""Latency of the drop-in AlignmentDecoder.decode (one utterance per call, the reference's usage
pattern, networks/task/forced_alignment.py:174-176) on BASELINE configs[0]: T=500, S=40, V=63."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hubertfa_b200 import synth
from hubertfa_b200.alignment_decoder import AlignmentDecoder

V = 63
vocab, items = synth.make_batch(np.array([500], np.int32), np.array([40], np.int32), V, seed=1234, planted=True)
it = items[0]
dev = torch.device("cuda")
head = torch.zeros(1, 500, V + 2, device=dev)
head[0, :, 0] = it["edge"][0].to(dev)
head[0, :, 2:] = it["frame"][0].to(dev)
frame, edge = head[:, :, 2:], head[:, :, 0]
ctc = torch.cat([head[:, :, [1]], head[:, :, 3:]], dim=-1)
dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
for _ in range(20):
    dec.decode(frame, edge, ctc, 9.99, it["ph_seq"], it["word_seq"], it["ph_idx_to_word_idx"])
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    out = dec.decode(frame, edge, ctc, 9.99, it["ph_seq"], it["word_seq"], it["ph_idx_to_word_idx"])
dt = (time.perf_counter() - t0) / n
print(f"decode(): {dt * 1e3:.3f} ms per utterance ({500 * 40 / dt:.3e} cells/s), {len(out[0])} phonemes")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(100):
    dec.decode(frame, edge, ctc, 9.99, it["ph_seq"], it["word_seq"], it["ph_idx_to_word_idx"])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
