"""GPU, tier 2: from logits through the drop-in ``AlignmentDecoder``.

The logits -> emission boundary is not bit-defined by the reference (torch's log_softmax / sigmoid
differ between CPU ISA paths and CUDA), so: emissions are compared element-wise with a tolerance
stated here, paths must match the reference on the golden set, and for any utterance of the random
sets whose path differs the oracle is re-run on OUR emissions and must then reproduce OUR path
exactly (which proves a <=few-ulp near-tie in third-party code, not a DP bug)."""
import numpy as np
import pytest
import torch

from conftest import golden_names
from gpu_util import bits, synth_core_inputs
from hubertfa_b200 import ops, synth
from hubertfa_b200.alignment_decoder import AlignmentDecoder
from oracle import c_oracle as oc
from oracle import hfa_oracle_np as onp

pytestmark = pytest.mark.gpu

EMIS_ATOL = 4e-6      # |log-softmax| values are O(1..20): a few f32 ulps of the max / log-sum terms
EDGE_ATOL = 2e-6      # on edge_log / not_edge_log away from the clamp (see test)


def _emission_gpu(frames, edges, ids_list, V):
    """Runs only the emission kernel; returns per-utterance (emis [T,S], edge2 [T,2], edge_p [T])."""
    dev = torch.device("cuda")
    T = [f.shape[0] for f in frames]
    S = [len(i) for i in ids_list]
    plan = ops.AlignPlan(T, S, np.concatenate(ids_list), V, 0.02)
    ws = plan.new_workspace(dev)
    plan.upload(ws)
    plan.set_inputs(ws, [f.data_ptr() for f in frames], [f.stride(0) for f in frames],
                    [f.stride(1) for f in frames], [e.data_ptr() for e in edges], [e.stride(0) for e in edges])
    ops.emission(ws, plan.handle, ops.TORCH_TO_DTYPE[frames[0].dtype])
    torch.cuda.synchronize()
    return plan, ws


def _ws_region(plan, ws, name):
    return plan.debug_region(ws, name)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("V", [39, 63, 74, 200, 300])
def test_emission_kernel_vs_oracle_and_torch(dtype, V):
    """Strided views of a [T, V+2] head output (networks/task/forced_alignment.py:288-291)."""
    dev = torch.device("cuda")
    shapes = [(130, 20), (64, 33), (257, 150), (1, 3), (63, 5), (65, 70)]
    heads, frames, edges, ids_list = [], [], [], []
    for i, (T, S) in enumerate(shapes):
        rng = np.random.default_rng(10 * V + i)
        vocab = synth.make_vocab(V)
        ph_seq, _, _ = synth.make_ph_seq(rng, S, V, "dictionary")
        ids_list.append(np.array([vocab["vocab"][p] for p in ph_seq], dtype=np.int32))
        g = torch.Generator().manual_seed(10 * V + i)
        head = (3.0 * torch.randn(1, T, V + 2, generator=g)).to(dtype).to(dev)
        heads.append(head)
        frames.append(head[0, :, 2:])
        edges.append(head[0, :, 0])
    plan, ws = _emission_gpu(frames, edges, ids_list, V)
    emis = _ws_region(plan, ws, "emis").cpu().numpy()
    edge2 = _ws_region(plan, ws, "edge2").cpu().numpy()
    edgep = _ws_region(plan, ws, "edge_p").cpu().numpy()
    eo = fo = 0
    for (T, S), head, ids in zip(shapes, heads, ids_list):
        Sp = (S + 3) // 4 * 4
        e = emis[eo:eo + T * Sp].reshape(T, Sp)
        x = head[0].float().cpu()
        # oracle (libm) and the reference's own torch call on CPU and on CUDA
        want_c = oc.emission(np.ascontiguousarray(x[:, 2:].numpy()), ids)
        want_t = onp.frame_log_probs(x[None, :, 2:], ids, V)[:, ids]
        want_cuda = onp.frame_log_probs(head[:, :, 2:], ids, V)[:, ids]
        for want in (want_c, want_t, want_cuda):
            np.testing.assert_allclose(e[:, :S], want, rtol=0, atol=EMIS_ATOL)
        assert np.isneginf(e[:, S:]).all()
        p = edgep[fo:fo + T]
        p_t = onp.edge_pred(x[None, :, 0])
        np.testing.assert_allclose(p, p_t, rtol=0, atol=2.5e-7)
        _, ep = onp.edge_streams(p)                      # exact from OUR p: isolates the f64 log
        el, ne = onp.edge_logs(ep)
        got = edge2[fo:fo + T]
        np.testing.assert_allclose(got[:, 0], el, rtol=1e-6, atol=EDGE_ATOL)
        np.testing.assert_allclose(got[:, 1], ne, rtol=1e-6, atol=EDGE_ATOL)
        eo += T * Sp
        fo += (T + 15) // 16 * 16


def test_emission_fallback_for_non_unit_column_stride():
    """Logits stored [V, T] (a transposed view: column stride != 1) cannot travel by TMA and take the
    thread-loaded block kernel; results must equal the TMA path on the same values."""
    dev = torch.device("cuda")
    V = 63
    shapes = [(200, 30), (65, 150), (64, 7)]
    ids_list, fr_t, fr_c, edges = [], [], [], []
    for i, (T, S) in enumerate(shapes):
        rng = np.random.default_rng(700 + i)
        vocab = synth.make_vocab(V)
        ph_seq, _, _ = synth.make_ph_seq(rng, S, V, "dictionary")
        ids_list.append(np.array([vocab["vocab"][p] for p in ph_seq], dtype=np.int32))
        g = torch.Generator().manual_seed(700 + i)
        x = (3.0 * torch.randn(V, T, generator=g)).to(dev)
        fr_t.append(x.t())                         # [T, V] view with strides (1, T)
        fr_c.append(x.t().contiguous())            # same values, unit column stride
        edges.append((2.0 * torch.randn(T, generator=g)).to(dev))
    outs = []
    for frames in (fr_t, fr_c):
        plan, ws = _emission_gpu(frames, edges, ids_list, V)
        outs.append((_ws_region(plan, ws, "emis").cpu().numpy().copy(),
                     _ws_region(plan, ws, "edge2").cpu().numpy().copy()))
    # same kept-id summation order in both kernels -> identical bits
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    fo = 0
    for (T, S) in shapes:
        assert np.array_equal(outs[0][1][fo:fo + T].view(np.uint32), outs[1][1][fo:fo + T].view(np.uint32))
        fo += (T + 15) // 16 * 16


@pytest.mark.parametrize("name", golden_names())
def test_decode_matches_reference_golden(golden, name):
    c = golden.case(name)
    m = c["meta"]
    dec = AlignmentDecoder(synth.make_vocab(m["V"]), {"hop_length": m["hop"], "sample_rate": m["sr"]})
    frame = torch.from_numpy(c["frame"])[None].cuda()
    edge = torch.from_numpy(c["edge"])[None].cuda()
    ctc = torch.zeros(1, frame.shape[1], m["V"]).cuda()
    ph, ph_iv, wd, wd_iv, conf = dec.decode(frame, edge, ctc, m["wav_length"], m["ph_seq"], m["word_seq"],
                                            m["ph_idx_to_word_idx"])
    assert np.array_equal(dec.ph_idx_seq, c["ph_idx_seq"])
    assert np.array_equal(dec.ph_time_int_pred, c["ph_time_int"])
    assert list(ph) == list(c["ph_seq_pred"]) and list(wd) == list(c["word_seq_pred"])
    assert ph_iv.shape == c["ph_intervals_pred"].shape and ph_iv.dtype == np.float64
    assert wd_iv.shape == c["word_intervals_pred"].shape
    np.testing.assert_allclose(ph_iv, c["ph_intervals_pred"], rtol=0, atol=1e-7)   # seconds
    np.testing.assert_allclose(wd_iv, c["word_intervals_pred"], rtol=0, atol=1e-7)
    assert isinstance(conf, np.float32)
    if np.isfinite(c["total_confidence"]):
        np.testing.assert_allclose(conf, c["total_confidence"], rtol=1e-4)
        fin = np.isfinite(c["frame_confidence"])
        np.testing.assert_allclose(dec.frame_confidence[fin], c["frame_confidence"][fin], rtol=2e-4, atol=1e-12)
    else:
        assert np.isnan(conf)
    np.testing.assert_allclose(dec.edge_prob, c["edge_prob"], atol=5e-7)
    assert dec.ph_frame_pred.shape == (c["prob_log"].shape[0], m["V"])
    assert dec.ph_frame_pred.dtype == np.float32
    if "ph_frame_pred" in c:                          # ad:56-59,73: the masked softmax, values this time
        np.testing.assert_allclose(dec.ph_frame_pred, c["ph_frame_pred"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(dec.ph_frame_pred.sum(axis=1), 1.0, rtol=1e-5)
    # plot() (ad:152-168, called by validation_step, forced_alignment.py:414): the seven arguments it hands to
    # tools.plot.plot_for_valid must be the reference's -- captured with a stand-in for the matplotlib function
    import sys
    import types
    got = {}
    fake = types.ModuleType("tools.plot")
    fake.plot_for_valid = lambda *a: got.setdefault("args", a) and "figure"
    saved = {k: sys.modules.get(k) for k in ("tools", "tools.plot")}
    sys.modules["tools"] = types.ModuleType("tools")
    sys.modules["tools.plot"] = fake
    try:
        T_used = dec.ph_frame_pred.shape[0]
        mel = torch.arange(8 * T_used, dtype=torch.float32).reshape(1, 8, T_used)
        assert dec.plot(mel.cuda()) == "figure"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    a = got["args"]
    assert len(a) == 7
    assert isinstance(a[0], np.ndarray) and np.array_equal(a[0], mel.numpy())              # melspec.cpu().numpy()
    assert list(a[1]) == list(c["plot_ph_seq"])                                             # ph_pred_seq
    assert np.array_equal(np.asarray(a[2]), c["plot_ph_intervals_int"]) and np.asarray(a[2]).dtype == np.int32
    fin = np.isfinite(c["frame_confidence"])
    np.testing.assert_allclose(a[3][fin], c["frame_confidence"][fin], rtol=2e-4, atol=1e-12)
    assert a[4].shape == (T_used, len(m["ph_seq"]))                                         # ph_frame_pred[:, ph_seq_id]
    np.testing.assert_allclose(np.asarray(a[4], dtype=np.float64).sum(axis=0), c["plot_ph_frame_prob_sum"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(np.asarray(a[5]), c["plot_ph_idx_frame"])                         # per-frame state index
    np.testing.assert_allclose(a[6], c["edge_prob"], atol=5e-7)


def test_decode_batch_c2_sized_vs_oracle():
    """BASELINE config 2 at full size: 256 utterances, 5-30 s, 20-150 phonemes, V=63."""
    V = 63
    T, S = synth.sample_shapes(256, seed=synth.SEED0)
    vocab, items = synth.make_batch(T, S, V, seed=synth.SEED0, planted=True)
    mel = synth.MELSPEC_50FPS
    dec = AlignmentDecoder(vocab, mel)
    frames = [it["frame"].cuda()[0] for it in items]
    edges = [it["edge"].cuda()[0] for it in items]
    res = dec.decode_batch(frames, edges, [it["ph_seq"] for it in items], [it["word_seq"] for it in items],
                           [it["ph_idx_to_word_idx"] for it in items])
    assert (res.status == 0).all()
    # our own emissions, for the explain-every-mismatch protocol
    plan, ws = _emission_gpu(frames, edges, [it["ids"] for it in items], V)
    emis = _ws_region(plan, ws, "emis").cpu().numpy()
    edge2 = _ws_region(plan, ws, "edge2").cpu().numpy()
    eo = fo = 0
    n_diff = 0
    for b, it in enumerate(items):
        t, s = int(T[b]), int(S[b])
        sp = (s + 3) // 4 * 4
        ours = np.ascontiguousarray(emis[eo:eo + t * sp].reshape(t, sp)[:, :s])
        el = np.ascontiguousarray(edge2[fo:fo + t, 0])
        ne = np.ascontiguousarray(edge2[fo:fo + t, 1])
        eo += t * sp
        fo += (t + 15) // 16 * 16
        out, ex = onp.decode(vocab, mel, it["frame"], it["edge"], None, None, it["ph_seq"], it["word_seq"],
                             it["ph_idx_to_word_idx"], full=True)
        idx, tim, iv = res.segments(b)
        # size-independent properties of any valid alignment
        assert tim[0] == 0 and (np.diff(tim) > 0).all() and (np.diff(idx) > 0).all()
        assert (np.diff(idx) <= 2).all() and idx[-1] >= len(it["ids"]) - 2
        # the DP itself: the oracle run on OUR emissions must give OUR path, always
        r = oc.decode(it["ids"], ours, el, ne)
        assert np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"]), b
        np.testing.assert_allclose(res.final_score[b], r["dp_path"][-1], rtol=1e-6)
        if not (np.array_equal(idx, ex["ph_idx_seq"]) and np.array_equal(tim, ex["ph_time_int"])):
            n_diff += 1          # a <=few-ulp near-tie in third-party log_softmax (explained above)
            continue
        got = res[b]
        assert list(got[0]) == list(out[0]) and list(got[2]) == list(out[2])
        np.testing.assert_allclose(got[1], out[1], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got[3], out[3], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got[4], out[4], rtol=1e-4)
    print(f"config 2: {n_diff} of 256 paths differ from the CPU-torch reference (all explained)")
    assert n_diff <= 2, f"{n_diff} of 256 paths differ from the oracle"


def test_decode_batch_equals_single_decode_and_packed_input():
    V = 39
    T, S = synth.sample_shapes(24, seed=5, min_s=1, max_s=6, s_lo=3, s_hi=60)
    vocab, items = synth.make_batch(T, S, V, seed=5)
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    frames = [it["frame"].cuda() for it in items]
    edges = [it["edge"].cuda() for it in items]
    seqs = [it["ph_seq"] for it in items]
    res = dec.decode_batch(frames, edges, seqs, [it["word_seq"] for it in items],
                           [it["ph_idx_to_word_idx"] for it in items], want_frame_confidence=True)
    packed = dec.decode_batch(torch.cat([f[0] for f in frames]), torch.cat([e[0] for e in edges]), seqs,
                              [it["word_seq"] for it in items], [it["ph_idx_to_word_idx"] for it in items],
                              lengths=[int(t) for t in T])
    for b, it in enumerate(items):
        one = dec.decode(frames[b], edges[b], it["ctc"].cuda(), None, it["ph_seq"], it["word_seq"],
                         it["ph_idx_to_word_idx"])
        for got in (res[b], packed[b]):
            assert list(got[0]) == list(one[0]) and list(got[2]) == list(one[2])
            assert np.array_equal(got[1].reshape(one[1].shape), one[1])
            assert np.array_equal(got[3].reshape(one[3].shape), one[3])
            assert bits(got[4]) == bits(one[4])
        f0, f1 = res.frame_off[b], res.frame_off[b + 1]
        assert np.array_equal(bits(res.frame_confidence[f0:f1]), bits(dec.frame_confidence))
        assert np.array_equal(dec.ctc(), onp.ctc_greedy(it["ctc"][0].numpy()))


def test_random_logits_tier2_protocol():
    """Unpeaked Gaussian logits (many near-ties): every mismatch must be explained by re-running the
    oracle DP on our own emissions."""
    V = 74
    T, S = synth.sample_shapes(96, seed=99, min_s=2, max_s=12, s_lo=5, s_hi=120)
    vocab, items = synth.make_batch(T, S, V, seed=99)
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    frames = [it["frame"].cuda()[0] for it in items]
    edges = [it["edge"].cuda()[0] for it in items]
    res = dec.decode_batch(frames, edges, [it["ph_seq"] for it in items])
    plan, ws = _emission_gpu(frames, edges, [it["ids"] for it in items], V)
    emis = _ws_region(plan, ws, "emis").cpu().numpy()
    edge2 = _ws_region(plan, ws, "edge2").cpu().numpy()
    eo = fo = 0
    differ = 0
    for b, it in enumerate(items):
        t, s = int(T[b]), int(S[b])
        sp = (s + 3) // 4 * 4
        ours = np.ascontiguousarray(emis[eo:eo + t * sp].reshape(t, sp)[:, :s])
        el = np.ascontiguousarray(edge2[fo:fo + t, 0])
        ne = np.ascontiguousarray(edge2[fo:fo + t, 1])
        eo += t * sp
        fo += (t + 15) // 16 * 16
        idx, tim, _ = res.segments(b)
        r = oc.decode(it["ids"], ours, el, ne)
        assert np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"]), \
            f"utterance {b}: DP on our own emissions does not reproduce our path"
        ex = onp.decode(vocab, synth.MELSPEC_50FPS, it["frame"], it["edge"], None, None, it["ph_seq"], full=True)[1]
        if not (np.array_equal(idx, ex["ph_idx_seq"]) and np.array_equal(tim, ex["ph_time_int"])):
            differ += 1
    print(f"tier-2: {differ} of {len(items)} paths differ from the CPU-torch reference (all explained)")
    assert differ <= 3


def test_host_pipeline_matches_decode_batch():
    """HostBatchAligner (chunked upload overlapped with the kernels) returns the same segments, in
    the caller's utterance order, as decode_batch on device-resident logits."""
    from hubertfa_b200.pipeline import BufferPool, HostBatchAligner
    V = 63
    T, S = synth.sample_shapes(40, seed=21, min_s=1, max_s=10, s_lo=4, s_hi=140)
    vocab, items = synth.make_batch(T, S, V, seed=21, planted=True)
    head = torch.cat([torch.cat([it["edge"][0][:, None], torch.zeros(int(t), 1), it["frame"][0]], dim=1)
                      for it, t in zip(items, T)]).pin_memory()
    ids_cat = np.concatenate([it["ids"] for it in items])
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    ref = dec.decode_batch([it["frame"].cuda() for it in items], [it["edge"].cuda() for it in items],
                           [it["ph_seq"] for it in items])
    pool = BufferPool(torch.device("cuda"))
    for n_chunks in (1, 3, 4, 64):
        for _ in range(2):                      # second pass reuses the pooled buffers
            al = HostBatchAligner(T, S, ids_cat, V, dec.frame_length, V + 2, n_chunks=n_chunks, pool=pool)
            out = al.run(head)
            assert out.all_ok() and (out["status"] == 0).all()
            for b in range(len(items)):
                idx, tim, iv = al.segments(out, b)
                ridx, rtim, riv = ref.segments(b)
                assert np.array_equal(idx, ridx) and np.array_equal(tim, rtim) and np.array_equal(iv, riv)
                assert bits(out["total_conf"][b]) == bits(ref.total_confidence[b])


def test_fused_forward_equals_three_stage_route(monkeypatch):
    """Small batches, halo-band routing (HFA_LAT_KERNEL=band): hfa_forward_fused (emissions computed by the DP kernel's producer warps, never
    written to HBM) must give the same bits as hfa_emission + hfa_viterbi_forward -- backpointers,
    kept dp, results -- for strided [T, V+2] head views, all band widths, S up to 1030."""
    import torch
    from hubertfa_b200 import _lib, ops, synth
    monkeypatch.setenv("HFA_LAT_KERNEL", "band")
    V = 63
    T = np.array([500, 130, 257, 700, 16, 17, 1, 333, 900, 1300], dtype=np.int32)
    S = np.array([40, 7, 65, 150, 5, 33, 3, 97, 256, 1030], dtype=np.int32)
    vocab, items = synth.make_batch(T, S, V, seed=4242, style="dictionary", planted=True)
    dev = torch.device("cuda")
    ids = np.concatenate([it["ids"] for it in items]).astype(np.int32)
    heads = []
    for it in items:                                  # the network head layout: col 0 edge, 2.. frame
        h = torch.zeros(it["frame"].shape[1], V + 2)
        h[:, 0] = it["edge"][0]
        h[:, 2:] = it["frame"][0]
        heads.append(h.to(dev))
    outs = []
    for fused in (False, True):
        plan = ops.AlignPlan(T, S, ids, V, 0.02)
        rt = plan.routing()
        assert rt["warp_utts"] == 0 and rt["cta_utts"] == 0 and rt["keeps_dp"]
        ws, res = plan.new_workspace(dev), plan.new_result(dev)
        plan.upload(ws)
        plan.set_inputs(ws, [h[:, 2:].data_ptr() for h in heads], [V + 2] * len(heads), [1] * len(heads),
                        [h[:, 0].data_ptr() for h in heads], [V + 2] * len(heads))
        if fused:
            ops.forward_fused(ws, plan.handle, _lib.DTYPE_F32)
        else:
            ops.emission(ws, plan.handle, _lib.DTYPE_F32)
            ops.viterbi_forward(ws, plan.handle, None)
        fc = torch.empty(plan.total_frames, dtype=torch.float32, device=dev)
        dpp = torch.empty(plan.total_frames, dtype=torch.float32, device=dev)
        ops.backtrace(ws, plan.handle, res, fc, dpp)
        torch.cuda.synchronize()
        outs.append((plan.debug_region(ws, "bp").cpu().numpy().copy(), res.cpu().numpy().copy(),
                     fc.cpu().numpy(), dpp.cpu().numpy()))
    for j in (0, 2, 3):                               # backpointer words, frame confidence, dp on the path
        assert np.array_equal(outs[0][j].view(np.uint8), outs[1][j].view(np.uint8))
    va, vb = plan.views(outs[0][1]), plan.views(outs[1][1])
    for key in ("status", "n_seg", "end_state", "final_score", "total_conf"):
        assert np.array_equal(va[key].view(np.uint8), vb[key].view(np.uint8)), key
    for b in range(plan.n_utt):                       # segment slots beyond n_seg are never written
        o, k = int(plan.seg_off[b]), int(va["n_seg"][b])
        for key in ("ph_idx_seq", "ph_time_int", "intervals"):
            assert np.array_equal(va[key][o:o + k], vb[key][o:o + k]), (b, key)
    # and decode_batch (hfa_align_batch picks the fused route on its own) agrees with it
    from hubertfa_b200.alignment_decoder import AlignmentDecoder
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    r = dec.decode_batch([h[None, :, 2:] for h in heads], [h[None, :, 0] for h in heads],
                         [it["ph_seq"] for it in items])
    v = ops.AlignPlan(T, S, ids, V, synth.MELSPEC_50FPS["hop_length"] / synth.MELSPEC_50FPS["sample_rate"]).views(outs[1][1])
    assert np.array_equal(r.n_seg, v["n_seg"])
    for b in range(len(items)):
        idx, tim, _ = r.segments(b)
        o, k = int(r.seg_off[b]), int(v["n_seg"][b])
        assert np.array_equal(idx, v["ph_idx_seq"][o:o + k]) and np.array_equal(tim, v["ph_time_int"][o:o + k])


def test_ctc_greedy_kernel_matches_numpy():
    """hfa_ctc_greedy vs the reference's three numpy lines (alignment_decoder.py:145-150): ties in the
    argmax (first maximum), strided views, half precisions, T = 1, all-blank, long T."""
    from hubertfa_b200 import ops
    g = torch.Generator().manual_seed(11)
    cases = []
    for T, V in [(1, 5), (7, 39), (300, 63), (1500, 74), (5000, 200), (33, 2)]:
        x = torch.randn(T, V, generator=g)
        cases.append(x)
        cases.append(torch.round(x * 2) / 2)                      # coarse grid: many exact ties
    cases.append(torch.zeros(40, 10))                             # all ties -> id 0 everywhere -> empty
    blank = torch.randn(50, 12, generator=g)
    blank[:, 0] += 100                                            # blank wins everywhere
    cases.append(blank)
    for x in cases:
        want = onp.ctc_greedy(x.numpy())
        got = ops.ctc_greedy(x.cuda()).cpu().numpy()
        assert np.array_equal(got, want)
        wide = torch.zeros(x.shape[0], 2 * x.shape[1] + 3)
        wide[:, 3::2] = x                                         # strided view of a wider tensor
        got = ops.ctc_greedy(wide.cuda()[:, 3::2]).cpu().numpy()
        assert np.array_equal(got, want)
    for dt in (torch.float16, torch.bfloat16):
        x = torch.randn(400, 63, generator=g).to(dt)
        assert np.array_equal(ops.ctc_greedy(x.cuda()).cpu().numpy(), onp.ctc_greedy(x.float().numpy()))


@pytest.mark.parametrize("lat_kernel", ["skew", "band"])
@pytest.mark.parametrize("mix", ["unsplit-fused", "split-bands", "with-long-sequences"])
def test_random_planted_batches_match_the_oracle_in_every_route(mix, lat_kernel, monkeypatch):
    """decode_batch (auto routing) on many small random utterances with peaked (planted) logits,
    head-layout strided views: paths, end states and intervals must equal the C oracle run on the
    same logits.  The three mixes exercise the fused forward pass (no utterance split), the
    multi-band route and the route with S > 256 bands."""
    monkeypatch.setenv("HFA_LAT_KERNEL", lat_kernel)
    rng = np.random.default_rng({"unsplit-fused": 1, "split-bands": 2, "with-long-sequences": 3}[mix])
    n = 120
    if mix == "unsplit-fused":
        S = rng.integers(1, 65, n)
    elif mix == "split-bands":
        S = rng.integers(1, 257, n)
    else:
        S = np.concatenate([rng.integers(1, 257, n - 6), rng.integers(257, 700, 6)])
    T = np.maximum(rng.integers(1, 400, n), (S * rng.uniform(0.6, 3.0, n)).astype(np.int64))
    T, S = T.astype(np.int32), S.astype(np.int32)
    V = 63
    vocab, items = synth.make_batch(T, S, V, seed=int(rng.integers(1 << 30)), style="dictionary", planted=True)
    dev = torch.device("cuda")
    heads = []
    for it in items:
        h = torch.zeros(1, it["frame"].shape[1], V + 2)
        h[0, :, 0] = it["edge"][0]
        h[0, :, 2:] = it["frame"][0]
        heads.append(h.to(dev))
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    res = dec.decode_batch([h[:, :, 2:] for h in heads], [h[:, :, 0] for h in heads],
                           [it["ph_seq"] for it in items], [it["word_seq"] for it in items],
                           [it["ph_idx_to_word_idx"] for it in items])
    bad = []
    for b, it in enumerate(items):
        e = oc.emission(it["frame"][0].numpy(), it["ids"])
        p = oc.edge_pred(it["edge"][0].numpy())
        _, ep = oc.edge_prob(p)
        el, ne = oc.edge_logs(ep)
        r = oc.decode(it["ids"], e, el, ne)
        idx, tim, iv = res.segments(b)
        if not (np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"])):
            bad.append(b)
            continue
        assert res.end_state[b] == r["end_state"]
        want = oc.intervals(int(T[b]), r["ph_time_int"], p, dec.frame_length)
        np.testing.assert_allclose(iv, want, rtol=0, atol=1e-7)
        if np.isfinite(r["dp_path"][-1]):
            assert abs(res.final_score[b] - r["dp_path"][-1]) <= 1e-4 * max(abs(r["dp_path"][-1]), 1.0)
    # planted logits are peaked; T < S utterances can still sit on exact near-ties of the two
    # softmax implementations -- those must be reproduced by the oracle DP on OUR emissions
    for b in bad:
        it = items[b]
        plan, ws = _emission_gpu([heads[b][0, :, 2:]], [heads[b][0, :, 0]], [it["ids"]], V)
        Sp = (int(S[b]) + 3) & ~3
        e_gpu = _ws_region(plan, ws, "emis").cpu().numpy().reshape(int(T[b]), Sp)[:, :int(S[b])]
        e2 = _ws_region(plan, ws, "edge2").cpu().numpy()[:int(T[b])]
        r = oc.decode(it["ids"], np.ascontiguousarray(e_gpu), np.ascontiguousarray(e2[:, 0]),
                      np.ascontiguousarray(e2[:, 1]))
        idx, tim, _ = res.segments(b)
        assert np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"]), (mix, b)
    assert len(bad) <= n // 10, f"{len(bad)} of {n} paths differ from the oracle on peaked logits"


def test_corpus_flow_sharded_and_chunked_equals_one_batch():
    """BASELINE configs[4] in miniature: a corpus sharded over two (virtual) ranks by cost, every shard
    streamed in chunks of bounded size, results merged in corpus order -- identical to aligning the
    whole corpus as one batch (which takes a different kernel route)."""
    from hubertfa_b200 import sharding
    V = 39
    T, S = synth.sample_shapes(360, seed=77, min_s=1, max_s=8, s_lo=3, s_hi=150)
    vocab, items = synth.make_batch(T, S, V, seed=77, planted=True)
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    frames = [it["frame"].cuda() for it in items]
    edges = [it["edge"].cuda() for it in items]
    whole = dec.decode_batch(frames, edges, [it["ph_seq"] for it in items])
    parts = []
    for shard in sharding.shard_by_cost(T, S, 2):
        segs = [None] * len(shard)
        conf = [None] * len(shard)
        for c in sharding.chunk_by_bytes(T[shard], S[shard], max_cells=1_500_000):
            idx = shard[c]
            r = dec.decode_batch([frames[i] for i in idx], [edges[i] for i in idx], [items[i]["ph_seq"] for i in idx])
            for j, pos in enumerate(c):
                segs[pos] = r.segments(j)
                conf[pos] = r.total_confidence[j]
        parts.append((shard, dict(segs=segs, conf=conf)))
    merged = sharding._merge(parts)
    assert len(merged["segs"]) == len(items) and all(s is not None for s in merged["segs"])
    for b in range(len(items)):
        idx, tim, iv = merged["segs"][b]
        widx, wtim, wiv = whole.segments(b)
        assert np.array_equal(idx, widx) and np.array_equal(tim, wtim) and np.array_equal(iv, wiv)
        np.testing.assert_allclose(merged["conf"][b], whole.total_confidence[b], rtol=1e-5)
