"""CPU: property tests tying the two oracles together on arbitrary inputs (hypothesis), so that the
C restatement the GPU tests compare against is checked far beyond the golden cases: random shapes,
random SP patterns, ties, -inf emissions, T < S."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as oc
from oracle import hfa_oracle_np as onp


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@st.composite
def core_inputs(draw):
    T = draw(st.integers(1, 40))
    S = draw(st.integers(1, 24))
    seed = draw(st.integers(0, 2**31 - 1))
    sp_rate = draw(st.sampled_from([0.0, 0.3, 0.5, 0.8, 1.0]))
    quant = draw(st.sampled_from([0, 1, 2]))          # 0: continuous, 1: coarse grid (ties), 2: constants
    rng = np.random.default_rng(seed)
    ids = np.where(rng.random(S) < sp_rate, 0, rng.integers(1, 30, S)).astype(np.int32)
    if quant == 0:
        e = (-rng.exponential(3.0, (T, S))).astype(np.float32)
        p = rng.random(T).astype(np.float32)
    elif quant == 1:
        e = (-rng.integers(0, 4, (T, S)) * 0.5).astype(np.float32)
        p = (rng.integers(0, 3, T) * 0.5).astype(np.float32)
    else:
        e = np.full((T, S), -1.0, np.float32)
        p = np.full(T, 0.25, np.float32)
    if draw(st.booleans()):
        e[rng.random((T, S)) < 0.05] = -np.inf
    return ids, e, p


@settings(max_examples=150, deadline=None)
@given(core_inputs())
def test_c_oracle_equals_numpy_oracle_bit_for_bit(x):
    ids, e, p = x
    _, ep = onp.edge_streams(p)
    el, ne = onp.edge_logs(ep)
    elc, nec = oc.edge_logs(ep)
    assert np.array_equal(_bits(el), _bits(elc)) and np.array_equal(_bits(ne), _bits(nec))
    dp, bt, _ = onp.forward_dp(ids, e, el, ne)
    r = oc.decode(ids, e, el, ne, full=True)
    assert r["rc"] == 0
    assert np.array_equal(_bits(r["dp"]), _bits(dp))
    assert np.array_equal(r["bt"], bt)
    idx, tim, dp_path, end_state = onp.backtrace(ids, dp, bt)
    assert np.array_equal(r["ph_idx_seq"], idx) and np.array_equal(r["ph_time_int"], tim)
    assert r["end_state"] == end_state
    assert np.array_equal(_bits(r["dp_path"]), _bits(dp_path))
    # structural invariants of any backtrace (alignment_decoder.py:274-280)
    assert tim[0] == 0 and (np.diff(tim) > 0).all() and (np.diff(idx) > 0).all() and (np.diff(idx) <= 2).all()
    # O(T) path-only rescoring == dp on the path (what the CUDA finalize kernel relies on)
    rs = onp.path_rescore(ids, e, el, ne, idx, tim)
    assert np.array_equal(_bits(rs), _bits(dp_path))
    # intervals: C == numpy, f64 bit-exact
    ed, _ = onp.edge_streams(p)
    iv = oc.intervals(e.shape[0], tim, p, 0.0116)
    assert np.array_equal(iv, onp.intervals_from_path(tim, ed, e.shape[0], 0.0116))
