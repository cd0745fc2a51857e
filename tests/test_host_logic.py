"""CPU: the C-ABI library loads and exports every symbol of include/hfa_align.h, the collation
("plan") logic, the batched SP-filter / word-merge, sharding, and the no-fallback guarantees."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from hubertfa_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hfa_align.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hfa_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in hfa_align.h but not exported"
    assert declared == set(_lib.SYMBOLS), "python binding and header disagree"
    assert _lib.load().hfa_abi_version() == _lib.ABI_VERSION


def test_plan_layout_and_statuses():
    from hubertfa_b200 import ops
    T = [500, 0, 30, 17, 40]
    S = [40, 3, 300, 0, 5]
    ids = np.zeros(sum(S), np.int32)
    ids[-1] = 99                      # out of range for V = 63 -> BAD_ID for the last utterance
    p = ops.AlignPlan(T, S, ids, 63, 0.02)
    assert list(p.seg_off) == [0, 40, 43, 343, 343, 348]
    assert list(p.frame_off) == [0, 500, 500, 530, 530, 530]
    assert p.total_cells == 500 * 40 + 30 * 300
    assert p.total_frames == 530
    L = p.layout
    offs = [L.status, L.n_seg, L.end_state, L.final_score, L.total_conf, L.ph_idx_seq, L.ph_time_int,
            L.intervals]
    assert offs == sorted(offs) and all(o % 16 == 0 for o in offs) and L.total_bytes >= offs[-1] + 16 * 348
    alg = p.algorithmic_bytes()
    # a batch this small goes to the banded kernel, whose forward pass also keeps dp (4 B per cell)
    assert p.routing()["keeps_dp"] and p.routing()["warp_utts"] == 0 and p.routing()["cta_utts"] == 0
    assert alg["dp"] == p.total_cells * 4 + p.total_frames * 8 + 4 * (32 * 40 + 2 * 300) + p.total_cells * 4
    assert p.algorithmic_bytes_fused() == p.total_frames * (63 * 4 + 4 + 12) + 4 * (32 * 40 + 2 * 300) \
        + p.total_cells * 4
    assert alg["emission"] == p.total_frames * (63 * 4 + 4) + p.total_cells * 4 + p.total_frames * 12
    p.close()
    with pytest.raises(Exception):
        ops.AlignPlan([1, 2], [1], np.zeros(1, np.int32), 63, 0.02)
    empty = ops.AlignPlan([], [], np.zeros(0, np.int32), 63, 0.02)
    assert empty.total_cells == 0 and empty.workspace_bytes >= 0


def test_pair_layout_is_chosen_for_the_batch(monkeypatch):
    """hfa_plan_create moves a batch to the warp kernel's SP-aware pair layout as a whole (every utterance without
    adjacent SPs and with at most 4 pairs per lane) when that lowers the frames-weighted instruction count
    (17 + 25 x pairs-per-lane against 16 + 21 x states-per-lane), else not at all."""
    from hubertfa_b200 import ops
    monkeypatch.setenv("HFA_LATENCY_MODE", "0")
    alt = lambda S: np.array([0 if i % 2 == 0 else 5 for i in range(S)], np.int32)     # SP ph SP ph ...
    dictionary = [alt(40), alt(30), alt(255),           # 20 / 15 / 128 pairs: 1 / 1 / 4 per lane vs 2 / 1 / 8 states
                  np.array([0, 0] + [4] * 60, np.int32),    # adjacent SPs: never
                  np.full(200, 3, np.int32)]                # 200 pairs, 7 per lane: never
    nosp = [np.full(50, 3, np.int32), np.full(90, 7, np.int32), alt(30)]   # pairs = states: plain is cheaper
    mk = lambda seqs: ops.AlignPlan([100] * len(seqs), [len(s) for s in seqs], np.concatenate(seqs), 63, 0.02)
    r = mk(dictionary).routing()
    assert r["warp_utts"] == 5 and r["pair_utts"] == 3
    assert mk(nosp).routing()["pair_utts"] == 0
    monkeypatch.setenv("HFA_PAIR", "2")
    assert mk(nosp).routing()["pair_utts"] == 3
    monkeypatch.setenv("HFA_PAIR", "0")
    assert mk(dictionary).routing()["pair_utts"] == 0
    monkeypatch.delenv("HFA_PAIR")
    big = ops.AlignPlan([100] * 3, [40, 30, 255], np.concatenate(dictionary[:3]), 300, 0.02)   # V > 255: plain rows
    assert big.routing()["pair_utts"] == 3
    assert big.routing()["stored_emission_bytes"] == 4 * 100 * (40 + 32 + 256)
    r = mk(dictionary).routing()                         # alt(): two distinct ids -> 4 columns instead of 40 / 32 / 256
    assert r["stored_emission_bytes"] == 4 * 100 * (4 + 4 + 4 + 64 + 200)


def test_compute_entry_points_fail_loudly_without_cuda():
    import torch
    from hubertfa_b200 import ops
    from hubertfa_b200._lib import HfaError
    from hubertfa_b200.alignment_decoder import AlignmentDecoder
    if torch.cuda.is_available():
        pytest.skip("this checks the CPU-only behaviour")
    dec = AlignmentDecoder({"vocab": {"SP": 0, "a": 1}, "vocab_size": 2}, {"hop_length": 512, "sample_rate": 44100})
    with pytest.raises((HfaError, RuntimeError)):
        dec.decode(torch.zeros(1, 5, 2), torch.zeros(1, 5), torch.zeros(1, 5, 2), None, ["SP", "a", "SP"])
    p = ops.AlignPlan([5], [3], np.array([0, 1, 0], np.int32), 2, 0.01)
    with pytest.raises((HfaError, RuntimeError, NotImplementedError)):
        ops.emission(torch.zeros(p.workspace_bytes, dtype=torch.uint8), p.handle, 0)


def test_decoder_raises_like_the_reference():
    from hubertfa_b200.alignment_decoder import AlignmentDecoder
    import torch
    dec = AlignmentDecoder({"vocab": {"SP": 0, "a": 1}, "vocab_size": 2}, {"hop_length": 512, "sample_rate": 44100})
    with pytest.raises(KeyError):                       # alignment_decoder.py:35
        dec.decode(torch.zeros(1, 5, 2), torch.zeros(1, 5), None, None, ["SP", "zz"])


def test_product_never_touches_the_oracle():
    """hubertfa_b200/ must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "hubertfa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "synth.py", f"{f} mentions the oracle"
    code = ("import sys; sys.path.insert(0, %r); import hubertfa_b200, hubertfa_b200.ops, "
            "hubertfa_b200.alignment_decoder, hubertfa_b200.sharding; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_batch_filter_and_merge_matches_the_reference_loop():
    """BatchAlignment's vectorised SP filter / word merge == the oracle's per-utterance loop
    (alignment_decoder.py:115-138) on fabricated raw segments (no GPU needed)."""
    from hubertfa_b200 import ops, synth
    from hubertfa_b200.alignment_decoder import BatchAlignment
    from oracle import hfa_oracle_np as onp
    rng = np.random.default_rng(3)
    n, V = 7, 39
    T = rng.integers(30, 90, n).astype(np.int32)
    S = rng.integers(1, 25, n).astype(np.int32)
    seqs = [synth.make_ph_seq(rng, int(s), V, ["dictionary", "alternate", "nosp"][i % 3]) for i, s in enumerate(S)]
    ids = np.concatenate([[0 if p == "SP" else int(p[1:]) for p in q[0]] for q in seqs]).astype(np.int32)
    plan = ops.AlignPlan(T, S, ids, V, 0.02)
    blob = np.zeros(plan.result_bytes, np.uint8)
    v = plan.views(blob)
    raw = []
    for b in range(n):
        k = int(rng.integers(1, S[b] + 1))
        idx = np.sort(rng.choice(S[b], size=k, replace=False))
        tim = np.sort(rng.choice(T[b], size=k, replace=False))
        tim[0] = 0
        times = np.concatenate([tim * 0.02 + rng.uniform(-0.01, 0.01, k), [T[b] * 0.02]])
        iv = np.stack([times[:-1], times[1:]], axis=1)
        o = int(plan.seg_off[b])
        v["n_seg"][b] = k
        v["ph_idx_seq"][o:o + k] = idx
        v["ph_time_int"][o:o + k] = tim
        v["intervals"][o:o + k] = iv
        v["total_conf"][b] = 0.5
        raw.append((idx, iv))
    is_sp = np.concatenate([np.array([p == "SP" for p in q[0]]) for q in seqs])
    widx = np.concatenate([np.asarray(q[2], dtype=np.int64) for q in seqs])
    res = BatchAlignment(plan, v, [q[0] for q in seqs], [q[1] for q in seqs], [q[2] for q in seqs], is_sp, widx)
    for b in range(n):
        want = onp.filter_and_merge(seqs[b][0], raw[b][0], raw[b][1], seqs[b][1], seqs[b][2])
        got = res[b]
        assert list(got[0]) == list(want[0]) and list(got[2]) == list(want[2])
        assert np.array_equal(got[1].reshape(want[1].shape), want[1])
        assert np.array_equal(got[3].reshape(want[3].shape), want[3])


def test_sharding_balances_cost_and_keeps_every_utterance():
    from hubertfa_b200 import sharding, synth
    T, S = synth.sample_shapes(1000, seed=1)
    for world in (1, 2, 4, 8):
        shards = sharding.shard_by_cost(T, S, world)
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(1000))
        cost = np.array([(T[s].astype(np.int64) * S[s]).sum() for s in shards])
        assert cost.max() <= 1.02 * cost.mean()
    chunks = sharding.chunk_by_bytes(T, S, max_cells=2_000_000)
    assert np.array_equal(np.concatenate(chunks), np.arange(1000))
    assert all((T[c].astype(np.int64) * S[c]).sum() <= 2_000_000 or len(c) == 1 for c in chunks)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from hubertfa_b200 import sharding, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    T, S = synth.sample_shapes(40, seed=9)
    mine = sharding.shard_by_cost(T, S, world)[rank]
    payload = {"n_seg": [int(S[i]) for i in mine], "tag": [f"utt{int(i)}@{rank}" for i in mine]}
    out = sharding.gather_on_host(mine, payload)
    if rank == 0:
        q.put((out["n_seg"], out["tag"]))
    dist.destroy_process_group()


def test_gather_on_host_world_size_2_gloo():
    import torch.multiprocessing as mp
    from hubertfa_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_seg, tag = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    T, S = synth.sample_shapes(40, seed=9)
    assert n_seg == [int(s) for s in S]
    assert all(t.startswith(f"utt{i}@") for i, t in enumerate(tag))


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """include/hfa_align.h must be usable from C (the drop-in boundary is a C ABI): compile a C
    translation unit with gcc -std=c99 -pedantic against it, link libhfa_align.so, run the host-only
    entry points (collation of a ragged batch) -- no GPU involved."""
    import os
    import shutil
    import subprocess
    from hubertfa_b200 import LIB_PATH
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "hfa_align.h"
int main(void) {
    const int32_t T[3] = {500, 0, 40}, S[3] = {40, 3, 5};
    int32_t ids[48] = {0};
    hfa_plan *plan = 0;
    HfaResultLayout lay;
    int32_t routing[8];
    if (hfa_abi_version() != HFA_ABI_VERSION) return 1;
    if (hfa_plan_create(3, 63, T, S, ids, 0.02, &plan) != HFA_OK) { puts(hfa_last_error()); return 2; }
    if (hfa_plan_total_cells(plan) != 500 * 40 + 40 * 5) return 3;
    if (hfa_plan_result_layout(plan, &lay) != HFA_OK || lay.total_bytes <= 0) return 4;
    if (hfa_plan_routing(plan, routing) != HFA_OK) return 5;
    {
        hfa_plan *bad = 0;
        if (hfa_plan_create(1, 0, T, S, ids, 0.02, &bad) != HFA_ERR_ARG || bad != 0) return 6;   /* bad vocabulary size */
    }
    printf("workspace bytes = %lld\n", (long long)hfa_plan_workspace_bytes(plan));
    hfa_plan_destroy(plan);
    return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-l:libhfa_align.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)


def test_corpus_plan_layout_mirrors_the_library():
    """hubertfa_b200.corpus reserves and parses chunk result blobs from (n_utt, sum S) alone -- every rank
    must derive the same offsets without talking to the others; it has to agree with hfa_plan_create."""
    from hubertfa_b200 import ops, synth
    from hubertfa_b200.corpus import CorpusPlan
    rng = np.random.default_rng(8)
    for n in (1, 2, 3, 7, 64, 257):
        T = rng.integers(1, 400, n).astype(np.int32)
        S = rng.integers(1, 90, n).astype(np.int32)
        ids = synth.make_ids_batch(T, S, 39, seed=n)
        plan = ops.AlignPlan(T, S, np.concatenate(ids), 39, 0.02)
        ns = int(S.sum())
        assert plan.result_bytes <= CorpusPlan.result_bytes(n, ns)
        a16 = lambda b: (b + 15) // 16 * 16
        o, L = 0, plan.layout
        for name, nb in (("status", 4 * n), ("n_seg", 4 * n), ("end_state", 4 * n), ("final_score", 4 * n),
                         ("total_conf", 4 * n), ("ph_idx_seq", 4 * ns), ("ph_time_int", 4 * ns), ("intervals", 16 * ns)):
            assert getattr(L, name) == o, name
            o += a16(nb)
    # sharding + chunking: every utterance in exactly one chunk of exactly one rank, blobs do not overlap
    T, S = synth.sample_shapes(3000, seed=4)
    cp = CorpusPlan(T, S, synth.make_ids_batch(T, S, 63, seed=4), 63, 0.02, 4, 20_000_000)
    seen = np.concatenate([np.concatenate(c) for c in cp.chunks])
    assert np.array_equal(np.sort(seen), np.arange(3000))
    offs = [o for r in cp.blob_off for o in r] + [cp.total_result_bytes]
    assert all(b > a for a, b in zip(offs[:-1], offs[1:]))
    loads = [cp.cells_of_rank(r) for r in range(4)]
    assert max(loads) <= 1.02 * min(loads)
    for r in range(4):
        sizes = [int((T[idx].astype(np.int64) * S[idx]).sum()) for idx in cp.chunks[r]]
        assert max(sizes) <= 1.1 * 20_000_000                      # the workspace bound (equal-sized chunks,
        assert max(sizes) <= 1.1 * min(sizes)                       # dealt round-robin: no small remainder chunk)


def _corpus_gather_worker(rank, world, port, name, q):
    import torch.distributed as dist
    from hubertfa_b200 import ops, synth
    from hubertfa_b200.corpus import CorpusPlan, SharedHostBuffer, all_status_ok, read_results
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    T, S = synth.sample_shapes(200, seed=12, min_s=1, max_s=4, s_lo=3, s_hi=40)
    ids = synth.make_ids_batch(T, S, 39, seed=12)
    cp = CorpusPlan(T, S, ids, 39, 0.02, world, 200_000)
    host = None
    if rank == 0:
        host = SharedHostBuffer(name, cp.total_result_bytes, create=True)
    dist.barrier()
    if rank != 0:
        host = SharedHostBuffer(name, cp.total_result_bytes, create=False)
    # every rank fabricates the result blobs of ITS chunks (what the D2H copies would deliver): utterance u gets
    # n_seg = 1 + u % 3 segments with ph_idx = u + j, time = 10 u + j
    for k, idx in enumerate(cp.chunks[rank]):
        plan = ops.AlignPlan(T[idx], S[idx], np.concatenate([ids[i] for i in idx]), 39, 0.02)
        blob = np.zeros(plan.result_bytes, np.uint8)
        v = plan.views(blob)
        for j, u in enumerate(idx):
            n = min(1 + int(u) % 3, int(S[u]))
            o = int(plan.seg_off[j])
            v["n_seg"][j] = n
            v["ph_idx_seq"][o:o + n] = int(u) + np.arange(n)
            v["ph_time_int"][o:o + n] = 10 * int(u) + np.arange(n)
            v["total_conf"][j] = float(u)
        o = cp.blob_off[rank][k]
        host.array[o:o + plan.result_bytes] = blob
    dist.barrier()                                   # the "gather": rank 0 now simply reads
    if rank == 0:
        ok = all_status_ok(cp, host)
        res = read_results(cp, host)
        good = ok and len(res) == 200
        for u in range(200):
            n = min(1 + u % 3, int(S[u]))
            good &= list(res[u]["ph_idx_seq"]) == [u + j for j in range(n)]
            good &= list(res[u]["ph_time_int"]) == [10 * u + j for j in range(n)] and res[u]["total_conf"] == float(u)
        q.put(bool(good))
    dist.barrier()
    host.close()
    dist.destroy_process_group()


def test_corpus_gather_through_shared_host_segment_world_size_2():
    """The multi-GPU gather path without GPUs: two ranks write their chunks' result blobs into the shared
    host segment at the offsets CorpusPlan derives, rank 0 reads everything back in corpus order."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    name = f"hfa_test_{os.getpid()}"
    procs = [ctx.Process(target=_corpus_gather_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
