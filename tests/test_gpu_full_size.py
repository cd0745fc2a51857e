"""GPU parity at the sizes BASELINE.json names (configs[2], [3], [4]) through the AUTO routing -- the
kernels a production call launches, not the ones an environment knob forces.

Protocol, per utterance:
  tier 1 (hard gate)  the C oracle DP (alignment_decoder.py:170-230, 264-283) run on OUR emissions
                      must reproduce OUR path and end state exactly and the final score to 1e-6;
  tier 2              the C oracle run from the LOGITS (libm softmax) must give the same path; every
                      utterance where it does not is counted, printed, and must be explained by
                      tier 1 (a few-ulp near-tie between two softmax implementations, not a DP bug).
"""
import numpy as np
import pytest
import torch

from hubertfa_b200 import ops, sharding, synth
from hubertfa_b200.alignment_decoder import AlignmentDecoder
from oracle import c_oracle as oc
from oracle import hfa_oracle_np as onp

pytestmark = pytest.mark.gpu


def planted_head(T, ids_list, V, seed):
    """Packed network-head output [sum T, V+2] f32 (col 0 edge, col 1 ctc blank, cols 2.. frame logits,
    networks/task/forced_alignment.py:288-291): Gaussian logits with a random monotone alignment of
    every utterance planted on top (+6 on its phoneme, +8 / -3 on the edge logit at / off a boundary)."""
    g = torch.Generator().manual_seed(int(seed))
    n_rows = int(np.sum(T))
    head = torch.randn(n_rows, V + 2, generator=g, dtype=torch.float32)
    head[:, 2:] *= 3.0
    head[:, 0] *= 2.0
    h = head.numpy()
    h[:, 0] -= 3.0
    rng = np.random.default_rng(int(seed))
    row0 = 0
    for t, ids in zip(T, ids_list):
        t, s = int(t), len(ids)
        if t >= s >= 1:
            cuts = np.sort(rng.choice(np.arange(1, t), size=s - 1, replace=False)) if s > 1 else np.zeros(0, np.int64)
            bounds = np.concatenate([[0], cuts, [t]]).astype(np.int64)
            state = np.repeat(np.arange(s), np.diff(bounds))
            h[row0 + np.arange(t), 2 + np.asarray(ids)[state]] += 6.0
            h[row0 + cuts, 0] += 8.0
        row0 += t
    return head


def emissions_of(dec, frames_dev, edges_dev, ids_list, V):
    """Our own emission kernel on the same logits -> host arrays for the tier-1 check."""
    dev = frames_dev[0].device
    T = [f.shape[0] for f in frames_dev]
    S = [len(i) for i in ids_list]
    plan = ops.AlignPlan(T, S, np.concatenate(ids_list), V, dec.frame_length)
    ws = plan.new_workspace(dev)
    plan.upload(ws)
    plan.set_inputs(ws, [f.data_ptr() for f in frames_dev], [f.stride(0) for f in frames_dev],
                    [f.stride(1) for f in frames_dev], [e.data_ptr() for e in edges_dev],
                    [e.stride(0) for e in edges_dev])
    ops.emission(ws, plan.handle, ops.TORCH_TO_DTYPE[frames_dev[0].dtype])
    torch.cuda.synchronize()
    emis = ops.unpack_emissions(plan, ws).cpu().numpy()     # dense [T][S] whatever the stored layout (compacted
    edge2 = plan.debug_region(ws, "edge2").cpu().numpy()    # rows for big batches in the pair layout)
    out, eo, fo = [], 0, 0
    for t, s in zip(T, S):
        out.append((np.ascontiguousarray(emis[eo:eo + t * s].reshape(t, s)),
                    np.ascontiguousarray(edge2[fo:fo + t, 0]), np.ascontiguousarray(edge2[fo:fo + t, 1])))
        eo += t * s
        fo += (t + 15) // 16 * 16
    return out


def check_batch(res, T, S, ids_list, head, V, frame_length, what, tier1_all=True, dec=None, head_dev=None):
    """res: BatchAlignment of the whole batch.  Returns the number of tier-2 mismatches."""
    hn = head.numpy()
    ids_cat = np.concatenate(ids_list)
    ref = oc.align_batch(T, S, V, hn[:, 2:], hn[:, 0], ids_cat, frame_length)
    assert ref["bad"] == 0
    row_off = np.concatenate([[0], np.cumsum(np.asarray(T, dtype=np.int64))])
    differ = []
    for b in range(len(T)):
        o, k = int(ref["seg_off"][b]), int(ref["n_seg"][b])
        idx, tim, iv = res.segments(b)
        same = (len(idx) == k and np.array_equal(idx, ref["ph_idx_seq"][o:o + k])
                and np.array_equal(tim, ref["ph_time_int"][o:o + k]))
        if not same:
            differ.append(b)
            continue
        np.testing.assert_allclose(iv.reshape(-1), ref["intervals"][2 * o:2 * (o + k)], rtol=0, atol=1e-7)
        np.testing.assert_allclose(res.total_confidence[b], ref["total_conf"][b], rtol=1e-4)
        # size-independent properties of any valid alignment
        assert tim[0] == 0 and (np.diff(tim) > 0).all() and (np.diff(idx) > 0).all() and (np.diff(idx) <= 2).all()
    todo = range(len(T)) if tier1_all else differ
    if len(list(todo)):
        sel = list(todo)
        frames = [head_dev[row_off[b]:row_off[b + 1], 2:] for b in sel]
        edges = [head_dev[row_off[b]:row_off[b + 1], 0] for b in sel]
        for b, (e, el, ne) in zip(sel, emissions_of(dec, frames, edges, [ids_list[b] for b in sel], V)):
            r = oc.decode(ids_list[b], e, el, ne)
            idx, tim, _ = res.segments(b)
            assert np.array_equal(idx, r["ph_idx_seq"]) and np.array_equal(tim, r["ph_time_int"]), \
                f"{what}: utterance {b} (T={T[b]}, S={S[b]}): the oracle DP on our emissions gives another path"
            assert res.end_state[b] == r["end_state"]
            np.testing.assert_allclose(res.final_score[b], r["dp_path"][-1], rtol=1e-6)
    print(f"{what}: {len(differ)} of {len(T)} paths differ from the libm-softmax oracle (all reproduced "
          f"by the oracle DP on our emissions)")
    return len(differ)


@pytest.mark.parametrize("V", [39, 74], ids=["japanese-V39", "jyutping-V74"])
def test_config4_4096_utterances_auto_routing(V):
    """BASELINE configs[3]: 4096 mixed-length utterances, 5-30 s, 20-150 phonemes, both vocabularies, one
    decode_batch call (auto routing: the one-warp-per-utterance kernels).  Tier 1 on EVERY utterance."""
    B = 4096
    T, S = synth.sample_shapes(B, seed=synth.SEED0 + V)
    ids_list = synth.make_ids_batch(T, S, V, seed=synth.SEED0 + V)
    head = planted_head(T, ids_list, V, seed=9000 + V)
    head_dev = head.cuda()
    vocab = synth.make_vocab(V)
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    names = np.array(["SP"] + [f"p{i}" for i in range(1, V)])
    ph_seqs = [list(names[i]) for i in ids_list]
    plan = ops.AlignPlan(T, S, np.concatenate(ids_list), V, dec.frame_length)
    rt = plan.routing()
    assert rt["warp_utts"] == B and rt["band_warps"] == 0, rt           # the big-batch route
    res = dec.decode_batch(head_dev[:, 2:], head_dev[:, 0], ph_seqs, lengths=[int(t) for t in T])
    assert (res.status == 0).all()
    n = check_batch(res, T, S, ids_list, head, V, dec.frame_length, f"config 4 (V={V})", True, dec, head_dev)
    assert n <= B // 200


def test_config3_from_logits_through_decode():
    """BASELINE configs[2]: T=30000, S=2000 from [1,T,V] logits through the drop-in decode() -- the
    emission kernel at that size, the long-sequence forward route, the table backtrace."""
    V, T, S = 63, 30000, 2000
    rng = np.random.default_rng(31337)
    vocab = synth.make_vocab(V)
    ph_seq, word_seq, ph2w = synth.make_ph_seq(rng, S, V, "dictionary")
    ids = np.array([vocab["vocab"][p] for p in ph_seq], dtype=np.int32)
    head = planted_head([T], [ids], V, seed=4711)
    head_dev = head.cuda()
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    ctc = torch.zeros(1, T, V, device="cuda")
    out = dec.decode(head_dev[None, :, 2:], head_dev[None, :, 0], ctc, None, ph_seq, word_seq, ph2w)
    # tier 1: the oracle DP on our emissions
    (e, el, ne), = emissions_of(dec, [head_dev[:, 2:]], [head_dev[:, 0]], [ids], V)
    r = oc.decode(ids, e, el, ne)
    assert np.array_equal(dec.ph_idx_seq, r["ph_idx_seq"]) and np.array_equal(dec.ph_time_int_pred, r["ph_time_int"])
    np.testing.assert_allclose(dec.final_score, r["dp_path"][-1], rtol=1e-6)
    fc, tot = oc.confidence(r["dp_path"])
    np.testing.assert_allclose(dec.frame_confidence, fc, rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out[4], tot, rtol=1e-4)
    # tier 2: the reference's own arithmetic (torch softmax on the CPU, numpy) from the logits
    want, ex = onp.decode(vocab, synth.MELSPEC_50FPS, head[None, :, 2:], head[None, :, 0], None, None,
                          ph_seq, word_seq, ph2w, full=True)
    assert np.array_equal(dec.ph_idx_seq, ex["ph_idx_seq"]) and np.array_equal(dec.ph_time_int_pred, ex["ph_time_int"])
    assert list(out[0]) == list(want[0]) and list(out[2]) == list(want[2])
    np.testing.assert_allclose(out[1], want[1], rtol=0, atol=1e-7)
    np.testing.assert_allclose(out[3], want[3], rtol=0, atol=1e-7)
    assert len(out[0]) > 500


def test_corpus_20000_utterances_sharded_chunked_auto_routing():
    """BASELINE configs[4] at a size the test budget allows: 20 000 utterances, sharded over two ranks by
    cost (shard_by_cost), every shard streamed in bounded chunks (chunk_by_bytes) through decode_batch
    (chunks are big batches: the one-warp-per-utterance route), merged in corpus order on the host --
    against the C oracle from the logits; every mismatch must pass tier 1."""
    V, B = 63, 20000
    T, S = synth.sample_shapes(B, seed=555, min_s=2, max_s=12, s_lo=10, s_hi=100)
    ids_list = synth.make_ids_batch(T, S, V, seed=555)
    head = planted_head(T, ids_list, V, seed=556)
    head_dev = head.cuda()
    row_off = np.concatenate([[0], np.cumsum(T.astype(np.int64))])
    vocab = synth.make_vocab(V)
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    names = np.array(["SP"] + [f"p{i}" for i in range(1, V)])
    parts, n_chunks = [], 0
    for shard in sharding.shard_by_cost(T, S, 2):
        segs, conf, fin, est = ([None] * len(shard) for _ in range(4))
        for c in sharding.chunk_by_bytes(T[shard], S[shard], max_cells=40_000_000):
            idx = shard[c]
            n_chunks += 1
            r = dec.decode_batch([head_dev[row_off[i]:row_off[i + 1], 2:] for i in idx],
                                 [head_dev[row_off[i]:row_off[i + 1], 0] for i in idx],
                                 [list(names[ids_list[i]]) for i in idx])
            assert (r.status == 0).all()
            for j, pos in enumerate(c):
                segs[pos], conf[pos], fin[pos], est[pos] = r.segments(j), r.total_confidence[j], r.final_score[j], r.end_state[j]
        parts.append((shard, dict(segs=segs, conf=conf, fin=fin, est=est)))
    merged = sharding._merge(parts)
    assert n_chunks >= 6 and all(x is not None for x in merged["segs"])

    class Merged:                                   # the BatchAlignment accessors check_batch uses
        total_confidence = np.array(merged["conf"], dtype=np.float32)
        final_score = np.array(merged["fin"], dtype=np.float32)
        end_state = np.array(merged["est"])

        @staticmethod
        def segments(b):
            return merged["segs"][b]

    n = check_batch(Merged, T, S, ids_list, head, V, dec.frame_length, f"corpus of {B} in {n_chunks} chunks",
                    False, dec, head_dev)
    assert n <= B // 200
