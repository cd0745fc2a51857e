"""GPU tests of the C-ABI contract itself (include/hfa_align.h): per-workspace input records, the
device-resident input table and CUDA-graph capture of a whole step, ownership of pinned results."""
import numpy as np
import pytest
import torch

from gpu_util import bits
from hubertfa_b200 import _lib, ops, synth
from hubertfa_b200.alignment_decoder import AlignmentDecoder
from oracle import c_oracle as oc

pytestmark = pytest.mark.gpu


def _batch(n, seed, V=63, planted=True, **kw):
    T, S = synth.sample_shapes(n, seed=seed, **kw)
    vocab, items = synth.make_batch(T, S, V, seed=seed, planted=planted)
    return T, S, vocab, items


def _head(items, V, dev):
    """Per-utterance [T, V+2] head outputs (col 0 edge, cols 2.. frame logits) on the device."""
    out = []
    for it in items:
        h = torch.zeros(it["frame"].shape[1], V + 2)
        h[:, 0] = it["edge"][0]
        h[:, 2:] = it["frame"][0]
        out.append(h.to(dev))
    return out


def _oracle_paths(items):
    out = []
    for it in items:
        e = oc.emission(it["frame"][0].numpy(), it["ids"])
        _, ep = oc.edge_prob(oc.edge_pred(it["edge"][0].numpy()))
        el, ne = oc.edge_logs(ep)
        out.append(oc.decode(it["ids"], e, el, ne))
    return out


def test_one_plan_two_workspaces_with_different_input_layouts():
    """hfa_set_inputs records the input layout per (plan, workspace): workspace A gets transposed logits
    (column stride != 1: the thread-loaded emission kernel), workspace B -- set AFTER A -- contiguous rows
    (the TMA-fed kernel).  Running A afterwards must still use A's layout (round 1 kept one flag on the
    plan: the last call won, and A would have been read with bulk row copies)."""
    V, dev = 63, torch.device("cuda")
    T, S, vocab, items = _batch(12, 77, V, min_s=1, max_s=6, s_lo=3, s_hi=70)
    ids = np.concatenate([it["ids"] for it in items])
    plan = ops.AlignPlan(T, S, ids, V, 0.02)
    wsA, wsB = plan.new_workspace(dev), plan.new_workspace(dev)
    resA, resB = plan.new_result(dev), plan.new_result(dev)
    fr_t = [it["frame"][0].t().contiguous().to(dev).t() for it in items]      # [T, V] views, strides (1, T)
    fr_c = [it["frame"][0].to(dev) for it in items]
    edges = [it["edge"][0].to(dev) for it in items]
    for ws, fr in ((wsA, fr_t), (wsB, fr_c)):
        plan.upload(ws)
        plan.set_inputs(ws, [f.data_ptr() for f in fr], [f.stride(0) for f in fr], [f.stride(1) for f in fr],
                        [e.data_ptr() for e in edges], [e.stride(0) for e in edges])
    ops.align_batch(wsA, plan.handle, _lib.DTYPE_F32, resA, None)
    ops.align_batch(wsB, plan.handle, _lib.DTYPE_F32, resB, None)
    torch.cuda.synchronize()
    a, b = plan.views(resA.cpu().numpy()), plan.views(resB.cpu().numpy())
    want = _oracle_paths(items)
    for v in (a, b):
        assert (v["status"] == 0).all()
        for i, r in enumerate(want):
            o, k = int(plan.seg_off[i]), int(v["n_seg"][i])
            assert np.array_equal(v["ph_idx_seq"][o:o + k], r["ph_idx_seq"])
            assert np.array_equal(v["ph_time_int"][o:o + k], r["ph_time_int"])
    assert np.array_equal(bits(a["final_score"]), bits(b["final_score"]))


@pytest.mark.parametrize("n_utt", [1, 48], ids=["one-utterance", "batch-of-48"])
def test_device_input_table_and_graph_replay(n_utt):
    """hfa_set_inputs_device + hfa_align_batch + the D2H copy captured into ONE CUDA graph (ops.GraphedStep):
    every replay equals the eager result, and rewriting the device table between replays points the same
    graph at other logits of the same shapes."""
    V, dev = 63, torch.device("cuda")
    if n_utt == 1:
        T, S = np.array([500], np.int32), np.array([40], np.int32)
        vocab, items = synth.make_batch(T, S, V, seed=5, planted=True)
    else:
        T, S, vocab, items = _batch(n_utt, 5, V, min_s=1, max_s=8, s_lo=3, s_hi=150)
    _, items2 = synth.make_batch(T, S, V, seed=5, planted=False)      # same phoneme sequences, other logits
    ids = np.concatenate([it["ids"] for it in items])
    plan = ops.AlignPlan(T, S, ids, V, 0.02)
    ws, res = plan.new_workspace(dev), plan.new_result(dev)
    host = torch.empty(plan.result_bytes, dtype=torch.uint8, pin_memory=True)
    plan.upload(ws)

    def table_for(heads):
        tab, stride = plan.input_table([h[:, 2:].data_ptr() for h in heads], [V + 2] * n_utt, [1] * n_utt,
                                       [h[:, 0].data_ptr() for h in heads], [V + 2] * n_utt)
        assert stride == V + 2
        return torch.from_numpy(tab.copy()), stride

    heads1, heads2 = _head(items, V, dev), _head(items2, V, dev)
    tab1, stride = table_for(heads1)
    tab2, _ = table_for(heads2)
    table = tab1.to(dev)
    step = ops.GraphedStep(plan, ws, res, host, _lib.DTYPE_F32, table=table, max_row_stride=stride)

    def eager(heads):
        ws2, res2 = plan.new_workspace(dev), plan.new_result(dev)
        plan.upload(ws2)
        plan.set_inputs(ws2, [h[:, 2:].data_ptr() for h in heads], [V + 2] * n_utt, [1] * n_utt,
                        [h[:, 0].data_ptr() for h in heads], [V + 2] * n_utt)
        ops.align_batch(ws2, plan.handle, _lib.DTYPE_F32, res2, None)
        torch.cuda.synchronize()
        return plan.views(res2.cpu().numpy())

    def same(got, want):
        for key in ("status", "n_seg", "end_state", "final_score", "total_conf"):
            assert np.array_equal(got[key].view(np.uint8), want[key].view(np.uint8)), key
        for b in range(n_utt):
            o, k = int(plan.seg_off[b]), int(want["n_seg"][b])
            for key in ("ph_idx_seq", "ph_time_int", "intervals"):
                assert np.array_equal(got[key][o:o + k], want[key][o:o + k]), (b, key)

    want1, want2 = eager(heads1), eager(heads2)
    assert not np.array_equal(want1["final_score"], want2["final_score"])
    for _ in range(3):
        host.zero_()
        step.replay()
        torch.cuda.synchronize()
        same(plan.views(host.numpy().copy()), want1)
    table.copy_(tab2.to(dev))                       # same graph, other logits
    step.replay()
    torch.cuda.synchronize()
    same(plan.views(host.numpy().copy()), want2)
    _lib.load().hfa_release_thread_resources()       # side streams are re-created on the next call
    step2 = ops.GraphedStep(plan, ws, res, host, _lib.DTYPE_F32, table=table, max_row_stride=stride)
    step2.replay()
    torch.cuda.synchronize()
    same(plan.views(host.numpy().copy()), want2)


def test_pipeline_result_survives_the_next_batch():
    """A PipelineResult owns its pinned result buffers: running another batch through the same BufferPool
    while the first result is still held must not change the first result."""
    from hubertfa_b200.pipeline import BufferPool, HostBatchAligner
    V = 63
    pool = BufferPool(torch.device("cuda"))
    runs = []
    for seed in (31, 32):
        T, S, vocab, items = _batch(30, seed, V, min_s=1, max_s=8, s_lo=4, s_hi=120)
        head = torch.cat([torch.cat([it["edge"][0][:, None], torch.zeros(int(t), 1), it["frame"][0]], dim=1)
                          for it, t in zip(items, T)]).pin_memory()
        ids_cat = np.concatenate([it["ids"] for it in items])
        al = HostBatchAligner(T, S, ids_cat, V, 0.02, V + 2, n_chunks=3, pool=pool)
        runs.append((al, al.run(head), _oracle_paths(items)))          # the first result stays alive
    for al, out, want in runs:                                          # read the FIRST one after the second ran
        assert out.all_ok()
        for b, r in enumerate(want):
            idx, tim, _ = al.segments(out, b)
            assert np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"]), b


def test_decode_direct_and_custom_op_dispatch_agree():
    """decode() reaches hfa_align_batch through ctypes by default (latency) or through the hfa::align_batch
    torch custom op (AlignmentDecoder.dispatch_through_torch_op): same entry point, same results, bit for bit."""
    V = 63
    T, S, vocab, items = _batch(6, 404, V, min_s=1, max_s=6, s_lo=3, s_hi=90)
    outs = []
    for through_op in (False, True):
        dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
        dec.dispatch_through_torch_op = through_op
        runs = []
        for it in items:
            r = dec.decode(it["frame"].cuda(), it["edge"].cuda(), it["ctc"].cuda(), None, it["ph_seq"], it["word_seq"],
                           it["ph_idx_to_word_idx"])
            runs.append((r, dec.ph_idx_seq.copy(), dec.ph_time_int_pred.copy(), dec.frame_confidence.copy()))
        res = dec.decode_batch([it["frame"].cuda() for it in items], [it["edge"].cuda() for it in items],
                               [it["ph_seq"] for it in items])
        outs.append((runs, res))
    for (ra, ia, ta, fa), (rb, ib, tb, fb) in zip(outs[0][0], outs[1][0]):
        assert list(ra[0]) == list(rb[0]) and list(ra[2]) == list(rb[2])
        assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[3], rb[3]) and bits(ra[4]) == bits(rb[4])
        assert np.array_equal(ia, ib) and np.array_equal(ta, tb) and np.array_equal(bits(fa), bits(fb))
    assert np.array_equal(outs[0][1].ph_idx_seq, outs[1][1].ph_idx_seq)
    assert np.array_equal(outs[0][1].raw_intervals, outs[1][1].raw_intervals)
