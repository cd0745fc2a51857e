"""CPU: the vectorised gap post-processing against the loop restatement in the
oracle and, where the reference tree exists, against tools/post_processing.py itself."""
import copy
import os
import sys
import time

import numpy as np
import pytest

from hubertfa_b200 import post_processing as pp
from oracle import hfa_oracle_np as onp
from oracle.reference_import import REFERENCE_ROOT, reference_available


def _random_tier(rng, n, wav_length):
    """n intervals with gaps of every interesting size, sometimes labelled AP, sometimes touching."""
    seq, rows = [], []
    t = float(rng.choice([0.0, 0.03, 0.12, 0.5]))
    for _ in range(n):
        dur = float(rng.uniform(0.02, 0.6))
        seq.append(str(rng.choice(["AP", "a", "b", "k", "AP", "n"])))
        rows.append([t, t + dur])
        t += dur + float(rng.choice([0.0, 0.0, 0.04, 0.09, 0.1, 0.15, 0.29, 0.3, 0.45]))
    iv = np.array(rows, dtype=np.float64)
    if n:
        scale = (wav_length - float(rng.choice([0.0, 0.05, 0.2]))) / max(iv[-1, 1], 1e-9)
        iv = iv * min(scale, 1.0)
    return seq, iv


def _batch(seed, n_utt):
    rng = np.random.default_rng(seed)
    seqs, ivs, wavs = [], [], []
    for _ in range(n_utt):
        n = int(rng.integers(0, 12))
        wav = float(rng.uniform(1.0, 8.0))
        s, iv = _random_tier(rng, n, wav)
        seqs.append(s); ivs.append(iv.reshape(-1, 2)); wavs.append(wav)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])])
    labels = [w for s in seqs for w in s]
    cat = np.concatenate(ivs) if len(labels) else np.zeros((0, 2))
    return seqs, ivs, wavs, off, labels, cat


@pytest.mark.parametrize("seed", range(6))
def test_batch_routines_equal_the_loop_restatement(seed):
    seqs, ivs, wavs, off, labels, cat = _batch(seed, 60)
    filled = pp.fill_small_gaps_batch(labels, cat, off, wavs)
    lab2, iv2, off2 = pp.add_sp_batch(labels, filled, off, wavs)
    for u, (s, iv, w) in enumerate(zip(seqs, ivs, wavs)):
        a, b = off[u], off[u + 1]
        if len(s):
            want = onp.fill_small_gaps_loop(s, iv, w)
            assert np.array_equal(filled[a:b], want), u
        else:
            want = iv
        wl, wiv = onp.add_sp_loop(s, want, w)
        assert lab2[off2[u]:off2[u + 1]] == wl, u
        assert np.array_equal(iv2[off2[u]:off2[u + 1]], wiv), u


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_against_the_reference_functions():
    sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    from tools import post_processing as ref          # numpy-only module of the reference
    for seed in range(4):
        seqs, ivs, wavs, off, labels, cat = _batch(100 + seed, 80)
        preds = [(f"utt{u}.wav", wavs[u], np.float32(0.5), np.array(seqs[u]), ivs[u].copy(),
                  np.array(seqs[u]), ivs[u].copy()) for u in range(len(seqs))]
        for p in preds:          # the decoder hands over shape-(0,) arrays for empty tiers (ad:135-138)
            pass
        want, want_log = ref.post_processing(copy.deepcopy(preds))
        got, got_log = pp.post_processing(copy.deepcopy(preds))
        assert [w[0] for w in want_log] == [g[0] for g in got_log]
        assert len(want) == len(got)
        for w, g in zip(want, got):
            assert w[0] == g[0] and w[1] == g[1]
            assert list(w[3]) == list(g[3]) and list(w[5]) == list(g[5])
            assert np.array_equal(np.asarray(w[4], dtype=np.float64), np.asarray(g[4], dtype=np.float64))
            assert np.array_equal(np.asarray(w[6], dtype=np.float64), np.asarray(g[6], dtype=np.float64))
        # the loop restatement in the oracle is pinned by the same reference run
        for u, (s, iv, wl) in enumerate(zip(seqs, ivs, wavs)):
            if len(s):
                _, r = ref.fill_small_gaps(np.array(s), iv.copy(), wl)
                assert np.array_equal(onp.fill_small_gaps_loop(s, iv, wl), r)
                rl, riv = ref.add_SP(np.array(s), r, wl)
                ol, oiv = onp.add_sp_loop(s, r, wl)
                assert list(rl) == ol and np.array_equal(np.asarray(riv, dtype=np.float64), oiv)


def test_batch_version_is_much_faster_than_the_loop():
    """Config-2-sized tier set (256 utterances x ~85 intervals): report both times."""
    rng = np.random.default_rng(0)
    seqs, ivs, wavs = [], [], []
    for _ in range(256):
        wav = float(rng.uniform(5.0, 30.0))
        s, iv = _random_tier(rng, int(rng.integers(20, 150)), wav)
        seqs.append(s); ivs.append(iv); wavs.append(wav)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])])
    labels = [w for s in seqs for w in s]
    cat = np.concatenate(ivs)
    t0 = time.perf_counter()
    filled = pp.fill_small_gaps_batch(labels, cat, off, wavs)
    pp.add_sp_batch(labels, filled, off, wavs)
    t_batch = time.perf_counter() - t0
    t0 = time.perf_counter()
    for s, iv, w in zip(seqs, ivs, wavs):
        onp.add_sp_loop(s, onp.fill_small_gaps_loop(s, iv, w), w)
    t_loop = time.perf_counter() - t0
    print(f"post-processing of {len(labels)} intervals: batch {1e3 * t_batch:.2f} ms, loop {1e3 * t_loop:.2f} ms")
    assert t_batch < t_loop
