"""ctypes binding of oracle/hfa_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhfa_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hfa_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.hfa_oracle_emission.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_int64, _i32p, _f32p, C.c_int64]
        L.hfa_oracle_emission.restype = None
        L.hfa_oracle_edge_pred.argtypes = [C.c_int32, C.c_void_p, C.c_int64, _f32p]
        L.hfa_oracle_edge_pred.restype = None
        L.hfa_oracle_edge_prob.argtypes = [C.c_int32, _f32p, _f64p, _f64p]
        L.hfa_oracle_edge_prob.restype = None
        L.hfa_oracle_edge_logs.argtypes = [C.c_int32, _f64p, _f32p, _f32p]
        L.hfa_oracle_edge_logs.restype = None
        L.hfa_oracle_decode.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, C.c_int64, _f32p, _f32p,
                                        _i32p, _i32p, _i32p, _f32p, _i32p, C.c_void_p, C.c_void_p]
        L.hfa_oracle_decode.restype = C.c_int
        L.hfa_oracle_confidence.argtypes = [C.c_int32, _f32p, C.c_void_p]
        L.hfa_oracle_confidence.restype = C.c_float
        L.hfa_oracle_intervals.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, C.c_double, _f64p]
        L.hfa_oracle_intervals.restype = None
        L.hfa_oracle_align_batch.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _i64p, _i64p,
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i32p,
                                             C.c_double, _i32p, _i32p, _i32p, _f64p, _f32p, _i32p,
                                             C.c_int32]
        L.hfa_oracle_align_batch.restype = C.c_int
        L.hfa_oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def emission(logits: np.ndarray, ph_ids: np.ndarray) -> np.ndarray:
    """logits f32 [T,V] (any strides) -> gathered log-probs f32 [T,S]."""
    assert logits.dtype == np.float32 and logits.ndim == 2
    T, V = logits.shape
    ids = np.ascontiguousarray(ph_ids, dtype=np.int32)
    out = np.empty((T, len(ids)), dtype=np.float32)
    st, sv = (s // 4 for s in logits.strides)
    lib().hfa_oracle_emission(T, V, len(ids), logits.ctypes.data, st, sv, ids, out, len(ids))
    return out


def edge_pred(edge_logits: np.ndarray) -> np.ndarray:
    assert edge_logits.dtype == np.float32 and edge_logits.ndim == 1
    p = np.empty(edge_logits.shape[0], dtype=np.float32)
    lib().hfa_oracle_edge_pred(len(p), edge_logits.ctypes.data, edge_logits.strides[0] // 4, p)
    return p


def edge_prob(p: np.ndarray):
    p = np.ascontiguousarray(p, dtype=np.float32)
    ep = np.empty(len(p), dtype=np.float64)
    ed = np.empty(len(p), dtype=np.float64)
    lib().hfa_oracle_edge_prob(len(p), p, ep, ed)
    return ed, ep


def edge_logs(edge_prob_: np.ndarray):
    ep = np.ascontiguousarray(edge_prob_, dtype=np.float64)
    el = np.empty(len(ep), dtype=np.float32)
    ne = np.empty(len(ep), dtype=np.float32)
    lib().hfa_oracle_edge_logs(len(ep), ep, el, ne)
    return el, ne


def decode(ph_ids, prob_log, edge_log, not_edge_log, full: bool = False):
    """Gathered emissions [T,S] + edge logs -> dict(ph_idx_seq, ph_time_int, dp_path, ...)."""
    ids = np.ascontiguousarray(ph_ids, dtype=np.int32)
    e = np.ascontiguousarray(prob_log, dtype=np.float32)
    T, S = e.shape
    idx = np.zeros(S, dtype=np.int32)
    tim = np.zeros(S, dtype=np.int32)
    n = np.zeros(1, dtype=np.int32)
    es = np.zeros(1, dtype=np.int32)
    dp_path = np.zeros(max(T, 1), dtype=np.float32)
    dp = np.empty((T, S), dtype=np.float32) if full else None
    bt = np.empty((T, S), dtype=np.int8) if full else None
    rc = lib().hfa_oracle_decode(T, S, ids, e, S, np.ascontiguousarray(edge_log, dtype=np.float32),
                                 np.ascontiguousarray(not_edge_log, dtype=np.float32), idx, tim, n,
                                 dp_path, es, dp.ctypes.data if full else None,
                                 bt.ctypes.data if full else None)
    return dict(rc=rc, ph_idx_seq=idx[: n[0]].copy(), ph_time_int=tim[: n[0]].copy(),
                dp_path=dp_path[:T], end_state=int(es[0]), dp=dp, bt=bt)


def confidence(dp_path: np.ndarray):
    d = np.ascontiguousarray(dp_path, dtype=np.float32)
    fc = np.empty(len(d), dtype=np.float32)
    tot = lib().hfa_oracle_confidence(len(d), d, fc.ctypes.data)
    return fc, np.float32(tot)


def intervals(T: int, ph_time_int, p, frame_length: float) -> np.ndarray:
    tim = np.ascontiguousarray(ph_time_int, dtype=np.int32)
    out = np.zeros((len(tim), 2), dtype=np.float64)
    lib().hfa_oracle_intervals(T, len(tim), tim, np.ascontiguousarray(p, dtype=np.float32),
                               float(frame_length), out.reshape(-1) if len(tim) else np.zeros(1))
    return out


def align_batch(T, S, V, frame_logits, edge_logits, ph_ids, frame_length, n_threads=0):
    """Ragged batch through the whole C path (pthread pool over utterances).

    frame_logits: f32 [sum T, V] (any row stride, unit column stride, e.g. head[:, 2:]),
    edge_logits: f32 [sum T] (any stride, e.g. head[:, 0]), ph_ids: concatenated [sum S]."""
    T = np.ascontiguousarray(T, dtype=np.int32)
    S = np.ascontiguousarray(S, dtype=np.int32)
    B = len(T)
    frame_logits = np.asarray(frame_logits)
    edge_logits = np.asarray(edge_logits)
    if frame_logits.ndim == 1:
        frame_logits = frame_logits.reshape(-1, V)
    assert frame_logits.dtype == np.float32 and edge_logits.dtype == np.float32
    assert frame_logits.strides[1] == 4 and frame_logits.shape[1] == V
    row_off = np.zeros(B, dtype=np.int64)
    seg_off = np.zeros(B, dtype=np.int64)
    if B > 1:
        row_off[1:] = np.cumsum(T[:-1].astype(np.int64))
        seg_off[1:] = np.cumsum(S[:-1].astype(np.int64))
    nseg_tot = int(S.astype(np.int64).sum())
    out = dict(n_seg=np.zeros(B, np.int32), ph_idx_seq=np.zeros(nseg_tot, np.int32),
               ph_time_int=np.zeros(nseg_tot, np.int32), intervals=np.zeros(2 * nseg_tot, np.float64),
               total_conf=np.zeros(B, np.float32), status=np.zeros(B, np.int32), seg_off=seg_off)
    bad = lib().hfa_oracle_align_batch(B, V, T, S, row_off, seg_off, frame_logits.ctypes.data,
                                       frame_logits.strides[0] // 4, edge_logits.ctypes.data,
                                       edge_logits.strides[0] // 4,
                                       np.ascontiguousarray(ph_ids, dtype=np.int32), float(frame_length),
                                       out["n_seg"], out["ph_idx_seq"], out["ph_time_int"],
                                       out["intervals"], out["total_conf"], out["status"], int(n_threads))
    out["bad"] = bad
    return out


def max_threads() -> int:
    return int(lib().hfa_oracle_max_threads())
