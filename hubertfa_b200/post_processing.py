"""Gap post-processing of the decoder's intervals -- SURVEY.md 8(f) rank 2, the step right after
the hot path (reference: tools/post_processing.py, called at infer.py:63).

Two entry points:

* ``post_processing(predictions, add_phone="SP")`` -- same signature, same result tuple layout and
  the same per-item try/except -> ``error_log`` behaviour as the reference (post_processing.py:68-105),
  so ``infer.py`` can import it instead.
* ``fill_small_gaps_batch`` / ``add_sp_batch`` -- the same rules applied to a whole ragged batch at
  once with NumPy (no Python loop over intervals), for the output of ``decode_batch``.

The rules are interval rewrites that only ever look at one gap (the end of interval i and the start
of interval i+1, post_processing.py:36-59), so they vectorise exactly: results are bit-identical to
the reference loop (tests/test_post_processing.py).
"""
from __future__ import annotations

import numpy as np

MIN_SP_LENGTH = 0.1      # post_processing.py:1
SP_MERGE_LENGTH = 0.3    # post_processing.py:2


def fill_small_gaps_batch(labels, intervals: np.ndarray, offsets: np.ndarray, wav_lengths) -> np.ndarray:
    """post_processing.py:31-65 for a ragged batch.

    labels: sequence of str, one per interval (all utterances concatenated); intervals: f64 [N,2];
    offsets: int [n_utt+1]; wav_lengths: float [n_utt].  Returns a new [N,2] array (the reference
    mutates in place; callers that want that can assign back)."""
    iv = np.array(intervals, dtype=np.float64, copy=True).reshape(-1, 2)
    offsets = np.asarray(offsets, dtype=np.int64)
    n = iv.shape[0]
    if n == 0:
        return iv
    wav = np.asarray(wav_lengths, dtype=np.float64)
    is_ap = np.fromiter((w == "AP" for w in labels), dtype=bool, count=n)
    nonempty = offsets[1:] > offsets[:-1]
    first = offsets[:-1][nonempty]
    last = offsets[1:][nonempty] - 1
    # :32-34 a short leading gap is absorbed by the first interval
    lead = (iv[first, 0] > 0) & (iv[first, 0] < MIN_SP_LENGTH)
    iv[first[lead], 0] = 0.0
    # :36-59 gaps between neighbours of the same utterance
    left = np.arange(n - 1)
    same = np.ones(n - 1, dtype=bool)
    same[last[last < n - 1]] = False                      # i is the last interval of its utterance
    end_l = iv[:-1, 1].copy()
    start_r = iv[1:, 0].copy()
    gap = start_r - end_l
    act = same & (end_l < start_r) & (gap < SP_MERGE_LENGTH)
    ap_l, ap_r = is_ap[:-1], is_ap[1:]
    mean = (end_l + start_r) / 2
    both = act & ap_l & ap_r                              # case 1: AP | AP -> meet in the middle
    only_l = act & ap_l & ~ap_r                           # case 2: AP on the left grows
    only_r = act & ~ap_l & ap_r                           # case 3: AP on the right grows
    none = act & ~ap_l & ~ap_r & (gap < MIN_SP_LENGTH)    # case 4: short gap, meet in the middle
    mid = both | none
    iv[left[mid], 1] = mean[mid]
    iv[left[mid] + 1, 0] = mean[mid]
    iv[left[only_l], 1] = start_r[only_l]
    iv[left[only_r] + 1, 0] = end_l[only_r]
    # :61-63 a short trailing gap is absorbed by the last interval
    w_last = wav[nonempty]
    tail = (iv[last, 1] < w_last) & (w_last - iv[last, 1] < MIN_SP_LENGTH)
    iv[last[tail], 1] = w_last[tail]
    return iv


def add_sp_batch(labels, intervals: np.ndarray, offsets: np.ndarray, wav_lengths, add_phone: str = "SP"):
    """post_processing.py:5-28 for a ragged batch.  Returns (labels list, intervals f64 [M,2],
    offsets [n_utt+1]) with the silence intervals inserted."""
    iv = np.asarray(intervals, dtype=np.float64).reshape(-1, 2)
    offsets = np.asarray(offsets, dtype=np.int64)
    wav = np.asarray(wav_lengths, dtype=np.float64)
    n_utt = len(offsets) - 1
    n = iv.shape[0]
    counts = np.diff(offsets)
    utt_of = np.repeat(np.arange(n_utt), counts)
    is_first = np.zeros(n, dtype=bool)
    is_first[offsets[:-1][counts > 0]] = True
    prev_end = np.empty(n, dtype=np.float64)
    prev_end[1:] = iv[:-1, 1]
    if n:
        prev_end[0] = 0.0
    # silence BEFORE interval i: the leading one [0, start] unless start <= 0 (:13-14,:24-26), or a
    # gap to the previous interval of the same utterance (:16-18)
    before = np.where(is_first, iv[:, 0] > 0, prev_end < iv[:, 0])
    before_start = np.where(is_first, 0.0, prev_end)
    # silence AFTER the last interval (:21-23), or the whole file for an empty utterance (:8-11)
    last = offsets[1:] - 1
    has = counts > 0
    after = np.zeros(n_utt, dtype=bool)
    after[has] = iv[last[has], 1] < wav[has]
    # output slots per utterance
    n_before = np.bincount(utt_of[before], minlength=n_utt) if n else np.zeros(n_utt, dtype=np.int64)
    out_counts = counts + n_before + after.astype(np.int64) + (~has).astype(np.int64)
    out_off = np.zeros(n_utt + 1, dtype=np.int64)
    np.cumsum(out_counts, out=out_off[1:])
    m = int(out_off[-1])
    out_iv = np.empty((m, 2), dtype=np.float64)
    out_lab = np.empty(m, dtype=object)
    if n:
        # position of interval i: its utterance's base + (#intervals and #silences before it)
        cum_before = np.cumsum(before) - np.repeat(np.concatenate([[0], np.cumsum(n_before)[:-1]]), counts)
        local = np.arange(n) - np.repeat(offsets[:-1], counts)
        pos = out_off[:-1][utt_of] + local + cum_before
        out_iv[pos] = iv
        out_lab[pos] = np.asarray(labels, dtype=object)
        sp_pos = pos[before] - 1
        out_iv[sp_pos, 0] = before_start[before]
        out_iv[sp_pos, 1] = iv[before, 0]
        out_lab[sp_pos] = add_phone
    tail_pos = out_off[1:][after] - 1
    out_iv[tail_pos, 0] = iv[last[after], 1]
    out_iv[tail_pos, 1] = wav[after]
    out_lab[tail_pos] = add_phone
    empty_pos = out_off[:-1][~has]
    out_iv[empty_pos, 0] = 0.0
    out_iv[empty_pos, 1] = wav[~has]
    out_lab[empty_pos] = add_phone
    return list(out_lab), out_iv, out_off


def post_processing(predictions, add_phone: str = "SP"):
    """Drop-in for tools/post_processing.py:68-105: list of
    (wav_path, wav_length, confidence, ph_seq, ph_intervals, word_seq, word_intervals) ->
    (res, error_log), every item processed with the batch routines above (one item at a time keeps
    the reference's per-item error isolation)."""
    res, error_log = [], []
    for wav_path, wav_length, confidence, ph_seq, ph_intervals, word_seq, word_intervals in predictions:
        try:
            out = []
            for seq, iv in ((ph_seq, ph_intervals), (word_seq, word_intervals)):
                iv = np.asarray(iv, dtype=np.float64)
                if iv.ndim != 2 or iv.shape[0] == 0:
                    # the reference indexes word_intervals[0, 0] first (:32): an empty tier raises
                    # IndexError there and the item lands in error_log
                    raise IndexError("index 0 is out of bounds for axis 0 with size 0")
                off = np.array([0, len(seq)], dtype=np.int64)
                filled = fill_small_gaps_batch(seq, iv, off, [wav_length])
                lab, new_iv, _ = add_sp_batch(list(seq), filled, off, [wav_length], add_phone)
                out.append((lab, [[float(a), float(b)] for a, b in new_iv]))
            res.append([wav_path, wav_length, confidence, out[0][0], out[0][1], out[1][0], out[1][1]])
        except Exception as e:  # noqa: BLE001  (same catch-all as the reference, :103-104)
            error_log.append([wav_path, e])
    return res, error_log


def post_process_batch(result, wav_lengths, add_phone: str = "SP"):
    """Gap filling + silence insertion for a ``BatchAlignment`` (decode_batch output).

    Returns a dict with ragged phoneme and word tiers: labels (list of str), intervals f64 [M,2] and
    offsets [n_utt+1] each.  Utterances that were not aligned (status) come out as one silence."""
    wav = np.asarray(wav_lengths, dtype=np.float64)
    out = {}
    ph_labels = [result._ph_seqs[u][i] for u, i in zip(result.ph_utt, result.ph_state)]
    word_labels = [result._word_seqs[u][i] for u, i in zip(result.word_utt, result.word_index)]
    for name, labels, iv, off in (("ph", ph_labels, result.ph_intervals, result.ph_off),
                                  ("word", word_labels, result.word_intervals, result.word_off)):
        filled = fill_small_gaps_batch(labels, iv, off, wav)
        lab, new_iv, new_off = add_sp_batch(labels, filled, off, wav, add_phone)
        out[name + "_seq"], out[name + "_intervals"], out[name + "_off"] = lab, new_iv, new_off
    return out
