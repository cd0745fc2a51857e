"""Synthetic inputs of the shapes BASELINE.json names.

There is no checkpoint and no dataset on the box, so every test and bench run uses logits
drawn here.  Logits are generated on the CPU from a seeded ``torch.Generator`` so that the CPU
oracle and the GPU path see identical bits; the caller copies them to the device.

Vocabulary layout follows the reference binarizer (binarize.py:79-87): ``SP`` -> 0, real phonemes
-> 1..V-1.  Phoneme sequences follow the reference G2P front-ends: ``dictionary`` mimics
networks/g2p/dictionary_g2p.py:16-42 (SP, then words of 1-2 phonemes each followed by SP),
``alternate`` mimics networks/g2p/phoneme_g2p.py:8-18 (ph/SP alternation), ``nosp`` has no SP.
"""
from __future__ import annotations

import numpy as np
import torch

VOCAB_SIZES = {"opencpop-extension": 63, "japanese_dict_full": 39, "jyutping": 74}
FRAME_SECONDS = 0.02  # BASELINE.json: "10 s ~ 500 frames"
MELSPEC_50FPS = {"hop_length": 882, "sample_rate": 44100}  # 882 / 44100 = 20 ms
SEED0 = 1234


def make_vocab(V: int) -> dict:
    vocab = {"SP": 0}
    for i in range(1, V):
        vocab[f"p{i}"] = i
    return {"vocab": vocab, "vocab_size": V}


def make_ph_seq(rng: np.random.Generator, S: int, V: int, style: str = "dictionary"):
    """Returns (ph_seq, word_seq, ph_idx_to_word_idx) with len(ph_seq) == S."""
    if S < 1:
        raise ValueError("S must be >= 1")
    ph_seq, word_seq, ph2w = [], [], []
    if style == "nosp":
        for i in range(S):
            ph_seq.append(f"p{int(rng.integers(1, V))}")
            word_seq.append(f"w{i}")
            ph2w.append(i)
        return ph_seq, word_seq, ph2w
    ph_seq.append("SP")
    ph2w.append(-1)
    w = 0
    while len(ph_seq) < S:
        n = 1 if style == "alternate" else int(rng.integers(1, 3))
        n = min(n, S - len(ph_seq))
        for _ in range(n):
            ph_seq.append(f"p{int(rng.integers(1, V))}")
            ph2w.append(w)
        word_seq.append(f"w{w}")
        w += 1
        if len(ph_seq) < S:
            ph_seq.append("SP")
            ph2w.append(-1)
    return ph_seq, word_seq, ph2w


def make_logits(seed: int, T: int, V: int, ids: np.ndarray | None = None, planted: bool = False):
    """CPU f32 tensors shaped like the network head's outputs: [1,T,V], [1,T], [1,T,V].

    planted=True adds +6 on the phoneme of a random monotone alignment of ``ids`` and +5 / -3 on
    the edge logit at / away from its boundaries, which gives the peaked paths a trained model has.
    """
    g = torch.Generator().manual_seed(int(seed))
    frame = 3.0 * torch.randn(1, T, V, generator=g, dtype=torch.float32)
    edge = 2.0 * torch.randn(1, T, generator=g, dtype=torch.float32)
    ctc = torch.randn(1, T, V, generator=g, dtype=torch.float32)
    if planted and ids is not None and T >= len(ids) >= 1:
        S = len(ids)
        cuts = np.sort(np.random.default_rng(int(seed)).choice(np.arange(1, T), size=S - 1,
                                                               replace=False)) if S > 1 else []
        bounds = np.concatenate([[0], cuts, [T]]).astype(np.int64)
        for k in range(S):
            frame[0, bounds[k]:bounds[k + 1], int(ids[k])] += 6.0
        edge -= 3.0
        for b in bounds[1:-1]:
            edge[0, b] += 8.0
    return frame, edge, ctc


def sample_shapes(B: int, seed: int = SEED0, min_s: int = 5, max_s: int = 30, s_lo: int = 20,
                  s_hi: int = 150):
    """Config C2/C4/C5 shapes: dur ~ U(min_s, max_s) s at 50 fps, S ~ U{s_lo..s_hi}, S <= T/3."""
    rng = np.random.default_rng(seed)
    dur = rng.uniform(min_s, max_s, size=B)
    T = np.rint(dur / FRAME_SECONDS).astype(np.int32)
    S = rng.integers(s_lo, s_hi + 1, size=B).astype(np.int32)
    S = np.minimum(S, np.maximum(T // 3, 1)).astype(np.int32)
    return T, S


def make_batch(T, S, V: int, seed: int = SEED0, style: str = "dictionary", planted: bool = False):
    """Per-utterance python lists: ph_seqs (str), word info, ids, and CPU logits."""
    rng = np.random.default_rng(seed)
    vocab = make_vocab(V)
    items = []
    for i, (t, s) in enumerate(zip(T, S)):
        ph_seq, word_seq, ph2w = make_ph_seq(rng, int(s), V, style)
        ids = np.array([vocab["vocab"][p] for p in ph_seq], dtype=np.int32)
        frame, edge, ctc = make_logits(seed + i, int(t), V, ids, planted)
        items.append(dict(ph_seq=ph_seq, word_seq=word_seq, ph_idx_to_word_idx=ph2w, ids=ids,
                          frame=frame, edge=edge, ctc=ctc))
    return vocab, items


def make_ids_batch(T, S, V: int, seed: int = SEED0, style: str = "dictionary"):
    """ids only (concatenated int32) -- for device-side logit generation at corpus scale."""
    rng = np.random.default_rng(seed)
    out = []
    for s in S:
        ph_seq, _, _ = make_ph_seq(rng, int(s), V, style)
        out.append(np.array([0 if p == "SP" else int(p[1:]) for p in ph_seq], dtype=np.int32))
    return out


def make_ids_corpus(S, V: int, seed: int = SEED0):
    """``make_ids_batch`` for corpus sizes (100 000 utterances): dictionary-style sequences -- SP, then words of
    1-2 phonemes each followed by SP, cut at S states (networks/g2p/dictionary_g2p.py:16-42) -- drawn with a few
    vectorised numpy calls per utterance instead of one per phoneme.  Deterministic in (S, V, seed)."""
    rng = np.random.default_rng(seed)
    S = np.asarray(S, dtype=np.int64)
    total = int(S.sum())
    # one stream of (SP + word) units long enough for every utterance; an utterance takes the next units
    wlen = rng.integers(1, 3, size=total + len(S)).astype(np.int64)          # phonemes per word: 1 or 2
    phon = rng.integers(1, V, size=2 * total + 2 * len(S)).astype(np.int32)
    out, wi, pi = [], 0, 0
    for s in S:
        s = int(s)
        n_units = s                                   # more than enough units for s states
        unit = wlen[wi:wi + n_units] + 1              # SP + word
        start = np.concatenate([[0], np.cumsum(unit)[:-1]])
        used = int(np.searchsorted(start, s, side="left"))
        ids = phon[pi:pi + s].copy()
        ids[start[:used]] = 0                         # the SP that opens every unit
        out.append(ids)
        wi += used
        pi += s
    return out
