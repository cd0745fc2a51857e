"""NumPy restatement of HubertFA's forced-alignment decoder -- TEST INFRASTRUCTURE ONLY.

Nothing under ``hubertfa_b200/`` may import this module.  It is the second, dependency-light
oracle (the first is ``oracle/hfa_oracle.c``): every array carries an explicit dtype so the
mixed f32/f64 rounding sequence of the reference is visible in the code.

Reference: ``tools/alignment_decoder.py`` (cited as ``ad:<line>`` below).
Parity status: PINNED -- ``tests/test_oracle_golden.py`` compares this file with the
unmodified reference wherever ``/root/reference`` exists, and ``tests/golden/*.npz`` holds
reference outputs for the machines where it does not.

The third-party arithmetic on the path (torch ``softmax``/``log_softmax``/``sigmoid``, ad:57,63,69)
is called through torch on CPU exactly as the reference does, so ``decode`` here reproduces the
reference bit for bit on CPU tensors.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64
NEG_INF32 = F32(-np.inf)


# --------------------------------------------------------------------------------------------
# stage 1/2: logits -> per-frame log-probabilities and the edge stream (ad:35-40, 53-71, 83-84)
# --------------------------------------------------------------------------------------------
def frame_log_probs(ph_frame_logits, ph_seq_id: np.ndarray, vocab_size: int) -> np.ndarray:
    """ad:37-40,53,62-65: log-softmax over {ids of ph_seq} U {0}; others pushed down by 1e9 (f32)."""
    import torch

    drop = np.ones(vocab_size, dtype=bool)
    drop[np.asarray(ph_seq_id)] = False
    drop[0] = False
    penalty = torch.from_numpy(drop).to(ph_frame_logits.device)[None, None, :] * 1e9  # f32
    x = ph_frame_logits.float() - penalty.float()
    return torch.log_softmax(x, dim=-1).squeeze(0).cpu().numpy().astype(F32)


def frame_probs(ph_frame_logits, ph_seq_id: np.ndarray, vocab_size: int) -> np.ndarray:
    """ad:56-59: the softmax twin of :func:`frame_log_probs` (only consumed by ``plot``)."""
    import torch

    drop = np.ones(vocab_size, dtype=bool)
    drop[np.asarray(ph_seq_id)] = False
    drop[0] = False
    penalty = torch.from_numpy(drop).to(ph_frame_logits.device)[None, None, :] * 1e9
    x = ph_frame_logits.float() - penalty.float()
    return torch.nn.functional.softmax(x, dim=-1).squeeze(0).cpu().numpy().astype(F32)


def edge_pred(ph_edge_logits) -> np.ndarray:
    """ad:68-71: clamp((sigmoid(x) - 0.1) / 0.8, 0, 1), f32 [T]."""
    import torch

    p = ((torch.sigmoid(ph_edge_logits.float()) - 0.1) / 0.8).clamp(0.0, 1.0)
    return p.squeeze(0).cpu().numpy().astype(F32)


def edge_streams(p: np.ndarray):
    """ad:83-84.  p is f32 [T]; both results are float64 (the f32 diff is promoted afterwards)."""
    p = np.asarray(p, dtype=F32)
    T = p.shape[0]
    edge_diff = np.zeros(T, dtype=F64)
    if T > 1:
        edge_diff[:-1] = (p[1:] - p[:-1]).astype(F32).astype(F64)
    shifted = np.zeros(T, dtype=F64)
    shifted[1:] = p[:-1].astype(F64)
    edge_prob = np.clip(p.astype(F64) + shifted, 0.0, 1.0)
    return edge_diff, edge_prob


def edge_logs(edge_prob: np.ndarray):
    """ad:241-242: f64 log, rounded to f32."""
    edge_prob = np.asarray(edge_prob, dtype=F64)
    return np.log(edge_prob + 1e-6).astype(F32), np.log(1 - edge_prob + 1e-6).astype(F32)


# --------------------------------------------------------------------------------------------
# stage 3: the DP (ad:170-230, 232-288), vectorised over states, serial over frames
# --------------------------------------------------------------------------------------------
def forward_dp(ph_seq_id: np.ndarray, prob_log: np.ndarray, edge_log: np.ndarray,
               not_edge_log: np.ndarray):
    """Returns (dp f32 [T,S], bt int8 [T,S] with row 0 == -1, curr f64 [S]).

    prob_log is the gathered [T,S] emission matrix (ad:239).
    """
    ids = np.asarray(ph_seq_id)
    prob_log = np.asarray(prob_log, dtype=F32)
    T, S = prob_log.shape
    if T < 1:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # ad:250
    dp = np.full((T, S), NEG_INF32, dtype=F32)           # ad:246
    bt = np.full((T, S), -1, dtype=np.int8)              # ad:247
    curr = np.full(S, -np.inf, dtype=F64)                # ad:245
    dp[0, 0] = prob_log[0, 0]
    curr[0] = prob_log[0, 0]
    if ids[0] == 0 and S > 1:                            # ad:252-254
        dp[0, 1] = prob_log[0, 1]
        curr[1] = prob_log[0, 1]
    ratio = F64(T) / F64(S)                              # ad:186 (int64 / int64 -> f64)
    is_sp = ids == 0
    # ad:191-202: state i may be entered from i-2 only across an id-0 state i-1
    jump_ok = np.zeros(S, dtype=bool)
    if S >= 3:
        jump_ok[2:] = is_sp[1:-1]
    with np.errstate(invalid="ignore"):
        for t in range(1, T):
            e = prob_log[t]
            base = (dp[t - 1] + e).astype(F32)
            stay = (base + not_edge_log[t]).astype(F32)                      # ad:177
            a = (base + edge_log[t]).astype(F32)                            # ad:183-185 (f32 part)
            adv = (a.astype(F64) + curr * ratio).astype(F32)                # ad:186 (f64 part)
            one = np.full(S, NEG_INF32, dtype=F32)
            one[1:] = adv[:-1]
            two = np.full(S, NEG_INF32, dtype=F32)
            if S >= 3:
                two[2:] = np.where(jump_ok[2:], adv[:-2], NEG_INF32)
            best = stay.copy()
            code = np.zeros(S, dtype=np.int8)
            m1 = one > best                                                  # ad:213-216 strict >
            best[m1] = one[m1]
            code[m1] = 1
            m2 = two > best
            best[m2] = two[m2]
            code[m2] = 2
            dp[t] = best
            bt[t] = code
            e64 = e.astype(F64)
            curr = np.where(code == 0, np.where(e64 > curr, e64, curr), e64)  # ad:220-224
            curr[is_sp] = 0.0                                                # ad:226-228
    return dp, bt, curr


def backtrace(ph_seq_id: np.ndarray, dp: np.ndarray, bt: np.ndarray):
    """ad:264-283.  Returns (ph_idx_seq int64, ph_time_int int64, dp_path f32 [T], end_state)."""
    ids = np.asarray(ph_seq_id)
    T, S = dp.shape
    s = S - 1
    if S >= 2 and dp[-1, -2] > dp[-1, -1] and ids[-1] == 0:  # ad:269
        s = S - 2
    end_state = s
    idx, tim = [], []
    dp_path = np.empty(T, dtype=F32)
    for t in range(T - 1, -1, -1):
        c = int(bt[t, s])
        assert c >= 0 or t == 0                              # ad:275
        dp_path[t] = dp[t, s]
        if c != 0:
            idx.append(s)
            tim.append(t)
            s -= c
    return (np.array(idx[::-1], dtype=np.int64), np.array(tim[::-1], dtype=np.int64), dp_path,
            end_state)


def frame_confidence_from_path(dp_path: np.ndarray) -> np.ndarray:
    """ad:284-288."""
    return np.exp(np.diff(np.pad(np.asarray(dp_path, dtype=F32), (1, 0), "constant",
                                 constant_values=0.0), 1))


def total_confidence(frame_confidence: np.ndarray):
    """ad:97."""
    with np.errstate(all="ignore"):
        return np.exp(np.mean(np.log(frame_confidence + 1e-6)) / 3)


def decode_core(ph_seq_id, ph_prob_log, edge_prob, full: bool = False):
    """The ``_decode`` boundary (ad:232-294): ids [S], log-probs [T,V] f32, edge_prob f64 [T]."""
    ids = np.asarray(ph_seq_id)
    prob_log = np.asarray(ph_prob_log, dtype=F32)[:, ids]   # ad:239
    el, ne = edge_logs(edge_prob)
    dp, bt, _ = forward_dp(ids, prob_log, el, ne)
    ph_idx_seq, ph_time_int, dp_path, end_state = backtrace(ids, dp, bt)
    with np.errstate(all="ignore"):
        fc = frame_confidence_from_path(dp_path)
    if full:
        return ph_idx_seq, ph_time_int, fc, dict(dp=dp, bt=bt, dp_path=dp_path,
                                                 end_state=end_state, prob_log=prob_log,
                                                 edge_log=el, not_edge_log=ne)
    return ph_idx_seq, ph_time_int, fc


def path_rescore(ph_seq_id, prob_log, edge_log, not_edge_log, ph_idx_seq, ph_time_int):
    """O(T) reconstruction of dp along a GIVEN path without the dp matrix.

    This is the algorithm the CUDA finalize kernel uses; it lives here so that the CPU tests can
    show it reproduces ``dp_path`` bit for bit against :func:`forward_dp`/:func:`backtrace`.
    """
    ids = np.asarray(ph_seq_id)
    T, S = prob_log.shape
    ratio = F64(T) / F64(S)
    state = np.empty(T, dtype=np.int64)
    bounds = list(ph_time_int) + [T]
    for k, s in enumerate(ph_idx_seq):
        state[bounds[k]:bounds[k + 1]] = s
    out = np.empty(T, dtype=F32)
    s0 = state[0]
    # ad:250-254: only state 0 (and state 1 behind a leading id-0 state) are seeded at t = 0;
    # a path that starts anywhere else is infeasible and carries -inf (SURVEY 8a edge cases).
    seeded = s0 == 0 or (s0 == 1 and ids[0] == 0 and S > 1)
    d = F32(prob_log[0, s0]) if seeded else NEG_INF32
    run_max = F64(d)                    # curr of the state the path is in (no SP zeroing at t=0)
    out[0] = d
    with np.errstate(invalid="ignore"):
        for t in range(1, T):
            s = state[t]
            sp = state[t - 1]
            if s == sp:
                d = F32(F32(d + prob_log[t, s]) + not_edge_log[t])
                run_max = max(run_max, F64(prob_log[t, s]))
            else:
                a = F32(F32(d + prob_log[t, sp]) + edge_log[t])
                d = F32(F64(a) + run_max * ratio)
                run_max = F64(prob_log[t, s])
            if ids[s] == 0:
                run_max = F64(0.0)
            out[t] = d
    return out


# --------------------------------------------------------------------------------------------
# stage 4/5: intervals, SP filter, word merge (ad:104-143)
# --------------------------------------------------------------------------------------------
def intervals_from_path(ph_time_int, edge_diff, T: int, frame_length: float) -> np.ndarray:
    """ad:104-113, f64 [K,2] (before the SP filter and before clip(min=0))."""
    ph_time_int = np.asarray(ph_time_int, dtype=np.int64)
    frac = np.clip(edge_diff[ph_time_int] / 2, -0.5, 0.5)
    times = frame_length * np.concatenate([ph_time_int.astype(F32).astype(F64) + frac, [F64(T)]])
    return np.stack([times[:-1], times[1:]], axis=1)


def filter_and_merge(ph_seq, ph_idx_seq, ph_intervals, word_seq, ph_idx_to_word_idx):
    """ad:115-138: drop "SP"-labelled segments, merge runs of one word, clip to >= 0."""
    keep = [k for k, i in enumerate(ph_idx_seq) if ph_seq[i] != "SP"]
    ph_out = [ph_seq[ph_idx_seq[k]] for k in keep]
    ph_iv = [ph_intervals[k, :] for k in keep]
    words, word_iv, last = [], [], -1
    for k in keep:
        w = ph_idx_to_word_idx[ph_idx_seq[k]]
        if w == last:
            word_iv[-1][1] = ph_intervals[k, 1]
        else:
            words.append(word_seq[w])
            word_iv.append([ph_intervals[k, 0], ph_intervals[k, 1]])
            last = w
    return (np.array(ph_out), np.array(ph_iv).clip(min=0, max=None), np.array(words),
            np.array(word_iv).clip(min=0, max=None))


def decode(vocab: dict, melspec_config: dict, ph_frame_logits, ph_edge_logits, ctc_logits,
           wav_length, ph_seq, word_seq=None, ph_idx_to_word_idx=None, full: bool = False):
    """The ``decode`` boundary (ad:26-143) from torch logits to the reference's 5-tuple."""
    ids = np.array([vocab["vocab"][ph] for ph in ph_seq])          # ad:35 (KeyError propagates)
    if word_seq is None:                                           # ad:41-43
        word_seq = ph_seq
        ph_idx_to_word_idx = np.arange(len(ph_seq))
    if wav_length is not None:                                     # ad:45-50
        n = int((wav_length * melspec_config["sample_rate"] + 0.5) / melspec_config["hop_length"])
        ph_frame_logits = ph_frame_logits[:, :n, :]
        ph_edge_logits = ph_edge_logits[:, :n]
    frame_length = melspec_config["hop_length"] / melspec_config["sample_rate"]
    log_probs = frame_log_probs(ph_frame_logits, ids, vocab["vocab_size"])
    p = edge_pred(ph_edge_logits)
    T = log_probs.shape[0]
    edge_diff, edge_prob = edge_streams(p)
    ph_idx_seq, ph_time_int, fc, extra = decode_core(ids, log_probs, edge_prob, full=True)
    conf = total_confidence(fc)
    iv = intervals_from_path(ph_time_int, edge_diff, T, frame_length)
    out = filter_and_merge(ph_seq, ph_idx_seq, iv, word_seq, ph_idx_to_word_idx) + (conf,)
    if full:
        extra.update(ph_seq_id=ids, log_probs=log_probs, edge_p=p, edge_diff=edge_diff,
                     edge_prob=edge_prob, ph_idx_seq=ph_idx_seq, ph_time_int=ph_time_int,
                     frame_confidence=fc, raw_intervals=iv)
        return out, extra
    return out


def ctc_greedy(ctc_logits_np: np.ndarray) -> np.ndarray:
    """ad:145-150: argmax, keep frames where the label changes and is not blank (0)."""
    lab = np.argmax(ctc_logits_np, axis=-1)
    prev = np.concatenate([[0], lab[:-1]])
    return lab[(lab != prev) & (lab != 0)]


# --------------------------------------------------------------------------------------------
# gap post-processing (tools/post_processing.py, cited as pp:<line>)
# --------------------------------------------------------------------------------------------
def fill_small_gaps_loop(seq, intervals, wav_length, min_sp=0.1, merge=0.3):
    """pp:31-65 restated as a plain loop over gaps; returns a new f64 [n,2] array."""
    iv = np.array(intervals, dtype=F64, copy=True)
    n = len(seq)
    if 0 < iv[0, 0] < min_sp:                                   # pp:32-34
        iv[0, 0] = 0
    for i in range(n - 1):                                      # pp:36-59
        a, b = iv[i, 1], iv[i + 1, 0]
        if not (a < b and b - a < merge):
            continue
        left_ap, right_ap = seq[i] == "AP", seq[i + 1] == "AP"
        if left_ap and not right_ap:
            iv[i, 1] = b
        elif right_ap and not left_ap:
            iv[i + 1, 0] = a
        elif (left_ap and right_ap) or b - a < min_sp:
            iv[i, 1] = iv[i + 1, 0] = (a + b) / 2
    if iv[-1, 1] < wav_length and wav_length - iv[-1, 1] < min_sp:   # pp:61-63
        iv[-1, 1] = wav_length
    return iv


def add_sp_loop(seq, intervals, wav_length, add_phone="SP"):
    """pp:5-28 restated: silence in front (unless the first interval starts at <= 0), in every gap,
    and at the end.  Returns (labels, f64 [m,2])."""
    if len(seq) == 0:
        return [add_phone], np.array([[0.0, wav_length]], dtype=F64)
    iv = np.asarray(intervals, dtype=F64)
    labels, rows = [], []
    cursor = 0.0
    for k, (w, (s, e)) in enumerate(zip(seq, iv)):
        if (k == 0 and s > 0) or (k > 0 and cursor < s):
            labels.append(add_phone)
            rows.append([cursor, s])
        labels.append(w)
        rows.append([s, e])
        cursor = e
    if cursor < wav_length:
        labels.append(add_phone)
        rows.append([cursor, wav_length])
    return labels, np.array(rows, dtype=F64)
