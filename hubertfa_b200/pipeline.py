"""Host-resident logits -> alignments with the upload overlapped with the kernels.

A batch that starts in (pinned) host memory is bound by the PCIe copy of the logits
(~55 GB/s measured on the B200 box), and a small batch is additionally bound by the serial chain of
its longest utterance.  ``HostBatchAligner`` cuts the batch into a few CONTIGUOUS chunks of
utterances, uploads them in order with one DMA each and aligns every chunk on its own stream as soon
as it has landed; the collation (``hfa_plan_create``) of chunk i+1 runs on the host while chunk i is
on the wire.  If the caller packs its batch longest utterance first (``longest_first_order`` -- the
length-bucketed collation of north_star item 4), the last chunk to arrive holds the shortest
utterances and only their short chain is left when the upload ends.

Measured alternatives that lost (tools/prof_pipeline.py, B200, config 2): reading the logits
straight out of pinned host memory from the emission kernel (zero-copy) reaches 37 GB/s against
55 GB/s for the DMA copy; uploading utterance by utterance in an arbitrary order costs ~8 us of
launch overhead per copy.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops

_STREAMS = {}


def _side_stream(device, i):
    """Process-wide side streams per device (creating streams per batch would be wasteful)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), i)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(device=device)
    return _STREAMS[key]


def longest_first_order(T) -> np.ndarray:
    """Permutation that packs a batch longest utterance first (stable)."""
    return np.argsort(-np.asarray(T, dtype=np.int64), kind="stable")


class BufferPool:
    """Reusable device / pinned-host byte buffers -- lets successive batches reuse their staging memory
    (pinned allocation is slow).  Device buffers are scratch: handed out in call order and rewound by
    ``reset()``.  Pinned buffers carry RESULTS: ``take_pinned`` checks one out and it stays out until the
    ``PipelineResult`` that owns it gives it back (``give_pinned``), so a result a caller still holds is
    never overwritten by a later batch."""

    def __init__(self, device):
        self.device = device
        self._dev, self._pin = [], []
        self._di = self._pi = 0
        self._pin_free = []

    def reset(self):
        self._di = self._pi = 0

    def take_pinned(self, nbytes):
        """A pinned buffer of at least nbytes that no live result owns (first fit, else a new one)."""
        for i, buf in enumerate(self._pin_free):
            if buf.numel() >= nbytes:
                return self._pin_free.pop(i)
        return torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, pin_memory=True)

    def give_pinned(self, buf):
        self._pin_free.append(buf)

    @staticmethod
    def _take(store, i, nbytes, make):
        if i < len(store) and store[i].numel() >= nbytes:
            return store[i][:nbytes]
        buf = make(max(int(nbytes * 1.25), 256))
        if i < len(store):
            store[i] = buf
        else:
            store.append(buf)
        return buf[:nbytes]

    def device_bytes(self, nbytes):
        t = self._take(self._dev, self._di, nbytes,
                       lambda n: torch.empty(n, dtype=torch.uint8, device=self.device))
        self._di += 1
        return t

    def pinned_bytes(self, nbytes):
        t = self._take(self._pin, self._pi, nbytes, lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True))
        self._pi += 1
        return t


class _Chunk:
    """Static description of one contiguous utterance range (everything that does not depend on the
    logits is prepared once, in HostBatchAligner.__init__)."""
    __slots__ = ("b0", "b1", "piece", "T", "S", "ids", "seg_off", "frame_ptr_off", "edge_ptr_off", "stride",
                 "ones", "stream")


class PipelineResult(dict):
    """What ``HostBatchAligner.run`` returns: a dict whose per-utterance arrays (``status``, ``n_seg``,
    ``total_conf``, ``final_score``) and per-chunk ragged views (``chunks``) are cut out of the pinned
    result blobs on first access -- the blobs are on the host when ``run`` returns, the Python-side
    slicing is not part of the critical path of a pipeline that only needs some of them."""

    def __init__(self, n_utt, live, lib, pool=None):
        super().__init__()
        self._n, self._live, self._lib, self._pool = n_utt, live, lib, pool
        self.d2h_bytes = sum(int(item[2].total_bytes) for item in live)

    def _materialise(self):
        if self._live is None:
            return
        n = self._n
        out = dict(status=np.empty(n, np.int32), n_seg=np.empty(n, np.int32), total_conf=np.empty(n, np.float32),
                   final_score=np.empty(n, np.float32), chunks=[])
        for c, h, lay, host_res, _, _ in self._live:
            # a private copy: the pinned buffer goes back to the pool below and later batches reuse it
            blob = host_res.numpy()[:int(lay.total_bytes)].copy()
            m, ns = c.b1 - c.b0, int(c.seg_off[-1])

            def v(off, dtype, count, blob=blob):
                return blob[off:off + count * np.dtype(dtype).itemsize].view(dtype)

            views = dict(status=v(lay.status, np.int32, m), n_seg=v(lay.n_seg, np.int32, m),
                         end_state=v(lay.end_state, np.int32, m), final_score=v(lay.final_score, np.float32, m),
                         total_conf=v(lay.total_conf, np.float32, m), ph_idx_seq=v(lay.ph_idx_seq, np.int32, ns),
                         ph_time_int=v(lay.ph_time_int, np.int32, ns),
                         intervals=v(lay.intervals, np.float64, 2 * ns).reshape(ns, 2))
            for k in ("status", "n_seg", "total_conf", "final_score"):
                out[k][c.b0:c.b1] = views[k]
            out["chunks"].append((c.b0, c.b1, c.seg_off, views))
        self._release()
        dict.update(self, out)

    def all_ok(self) -> bool:
        """True when every utterance was aligned (status 0) -- reads the status words straight out
        of the pinned result blobs, without building the per-utterance arrays."""
        if self._live is None:
            return bool((dict.__getitem__(self, "status") == 0).all())
        for c, _, lay, host_res, _, _ in self._live:
            m = c.b1 - c.b0
            if host_res.numpy()[lay.status:lay.status + 4 * m].view(np.int32).any():
                return False
        return True

    def _release(self):
        """Destroys the chunk plans and hands the pinned result buffers back to the pool."""
        if self._live is not None:
            for item in self._live:
                self._lib.hfa_plan_destroy(item[1])
                if self._pool is not None:
                    self._pool.give_pinned(item[3])
            self._live = None

    def __getitem__(self, key):
        if not dict.__contains__(self, key):
            self._materialise()
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        self._materialise()
        return dict.__contains__(self, key)

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class HostBatchAligner:
    """One ragged batch whose head outputs live in pinned host memory as a packed [sum T, W] tensor.

    T, S: per-utterance frame / state counts (utterance b owns rows sum(T[:b]) .. +T[b]);
    ph_ids: concatenated phoneme ids; frame_col / edge_col: first column of the V frame logits and
    the column of the edge logit inside a row (networks/task/forced_alignment.py:288-291: 2 and 0).
    The hot path (``run``) talks to the C ABI directly through ctypes: per-chunk host overhead is
    what bounds the pipeline depth.  The upload is cut by ROWS (byte shares of the packed tensor), so
    its first piece is on the wire before any per-utterance table exists; utterance chunk c holds the
    utterances that are complete once piece c has landed.
    """

    # byte shares of the pieces, first to last: the early ones are big (their alignment overlaps the
    # rest of the upload anyway), the last one is small so that little is left when the upload ends
    DEFAULT_SHARES = (0.30, 0.28, 0.22, 0.12, 0.08)

    def __init__(self, T, S, ph_ids, vocab_size, frame_length, row_width, frame_col=2, edge_col=0,
                 n_chunks=None, device=None, dtype=torch.float32, pool: "BufferPool | None" = None,
                 shares=None):
        self.lib = _lib.load()
        self.T = np.ascontiguousarray(T, dtype=np.int32)
        self.S = np.ascontiguousarray(S, dtype=np.int32)
        self.ids = np.ascontiguousarray(ph_ids, dtype=np.int32)
        self.n_utt = int(self.T.size)
        self.vocab_size, self.frame_length = int(vocab_size), float(frame_length)
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        self.dt = ops.TORCH_TO_DTYPE[dtype]
        self.row_width = int(row_width)
        self.frame_col, self.edge_col = int(frame_col), int(edge_col)
        self.esz = torch.empty(0, dtype=dtype).element_size()
        if shares is None:
            shares = self.DEFAULT_SHARES if n_chunks is None else [1.0 / max(int(n_chunks), 1)] * max(int(n_chunks), 1)
        self.shares = np.asarray(shares, dtype=np.float64) / float(np.sum(shares))
        self.total_rows = int(self.T.sum(dtype=np.int64))
        # rows at which the upload is cut (the last piece ends at total_rows)
        cuts = np.minimum((self.total_rows * np.cumsum(self.shares)).astype(np.int64), self.total_rows)
        cuts[-1] = self.total_rows
        self.row_cuts = [int(x) for x in cuts]
        self.chunks = None
        self.pool = pool if pool is not None else BufferPool(self.dev)
        self.h2d_bytes = self.total_rows * self.row_width * self.esz
        self.d2h_bytes = 0

    def _build_chunks(self):
        W, esz = self.row_width, self.esz
        row_off = np.concatenate([[0], np.cumsum(self.T.astype(np.int64))])
        seg_off = np.concatenate([[0], np.cumsum(np.maximum(self.S, 0).astype(np.int64))])
        # utterance b is complete when the piece that holds its last row has landed
        ends = np.searchsorted(row_off[1:], np.asarray(self.row_cuts, dtype=np.int64), side="right")
        ends[-1] = self.n_utt
        self.chunks = []
        b0 = 0
        for ci, b1 in enumerate(int(e) for e in ends):
            if b1 <= b0:
                continue
            c = _Chunk()
            c.b0, c.b1, c.piece = b0, b1, ci
            c.T = np.ascontiguousarray(self.T[b0:b1])
            c.S = np.ascontiguousarray(self.S[b0:b1])
            c.ids = np.ascontiguousarray(self.ids[seg_off[b0]:seg_off[b1]])
            c.seg_off = seg_off[b0:b1 + 1] - seg_off[b0]
            rows = row_off[b0:b1]
            c.frame_ptr_off = np.ascontiguousarray((rows * W + self.frame_col) * esz)
            c.edge_ptr_off = np.ascontiguousarray((rows * W + self.edge_col) * esz)
            c.stride = np.full(b1 - b0, W, dtype=np.int64)
            c.ones = np.ones(b1 - b0, dtype=np.int64)
            c.stream = _side_stream(self.dev, 1 + ci)
            self.chunks.append(c)
            b0 = b1

    def run(self, head_host: torch.Tensor, profile: bool = False) -> dict:
        """head_host: pinned host tensor [sum T, row_width].  Returns per-utterance arrays
        (status, n_seg, total_conf, final_score) plus the per-chunk ragged views."""
        if head_host.is_cuda or head_host.dtype != self.dtype or head_host.dim() != 2 \
                or head_host.shape[1] != self.row_width or head_host.shape[0] < self.total_rows:
            raise ValueError("head_host must be a host tensor [sum T, row_width] of the configured dtype")
        import ctypes as C
        import time
        lib, W, esz, pool = self.lib, self.row_width, self.esz, self.pool
        pool.reset()
        live = []
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream()
            copy_stream = _side_stream(self.dev, 0)
            copy_stream.wait_stream(cur)
            t_host0 = time.perf_counter()
            if profile:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(cur)
            # 1. every piece goes on the wire, in order, before any host-side collation
            dev_head = pool.device_bytes(max(self.total_rows, 1) * W * esz).view(self.dtype).view(-1, W)
            landed = []
            with torch.cuda.stream(copy_stream):
                r0 = 0
                for r1 in self.row_cuts:
                    if r1 > r0:
                        dev_head[r0:r1].copy_(head_host[r0:r1], non_blocking=True)
                    ev = torch.cuda.Event(enable_timing=profile)
                    ev.record(copy_stream)
                    landed.append(ev)
                    r0 = r1
            if self.chunks is None:
                self._build_chunks()
            base = dev_head.data_ptr()
            # 2. collation + launches of chunk i while the later pieces are still travelling
            prof = []
            for c in self.chunks:
                n = c.b1 - c.b0
                h = C.c_void_p()
                _lib.check(lib.hfa_plan_create(n, self.vocab_size, c.T.ctypes.data, c.S.ctypes.data,
                                               c.ids.ctypes.data, self.frame_length, C.byref(h)), "hfa_plan_create")
                lay = _lib.ResultLayout()
                lib.hfa_plan_result_layout(h, C.byref(lay))
                ws = pool.device_bytes(max(int(lib.hfa_plan_workspace_bytes(h)), 256))
                res = pool.device_bytes(int(lay.total_bytes))
                host_res = pool.take_pinned(int(lay.total_bytes))     # owned by the result until it lets go
                st = c.stream
                st.wait_stream(cur)
                sp = int(st.cuda_stream)
                fp, ep = c.frame_ptr_off + base, c.edge_ptr_off + base
                wp = ws.data_ptr()
                _lib.check(lib.hfa_plan_upload(h, wp, sp), "hfa_plan_upload")
                _lib.check(lib.hfa_set_inputs(h, wp, fp.ctypes.data, c.stride.ctypes.data, c.ones.ctypes.data,
                                              ep.ctypes.data, c.stride.ctypes.data, sp), "hfa_set_inputs")
                st.wait_event(landed[c.piece])
                _lib.check(lib.hfa_align_batch(h, wp, self.dt, res.data_ptr(), None, sp), "hfa_align_batch")
                with torch.cuda.stream(st):
                    host_res[:int(lay.total_bytes)].copy_(res, non_blocking=True)
                    if profile:
                        done = torch.cuda.Event(enable_timing=True)
                        done.record(st)
                        prof.append((landed[c.piece], done, time.perf_counter() - t_host0))
                live.append((c, h, lay, host_res, ws, res))
            for item in live:
                item[0].stream.synchronize()
        out = PipelineResult(self.n_utt, live, lib, pool)
        self.d2h_bytes = out.d2h_bytes
        if profile:
            dict.__setitem__(out, "profile", [dict(landed_ms=ev0.elapsed_time(l), done_ms=ev0.elapsed_time(d),
                                                   host_issued_ms=1e3 * t) for l, d, t in prof])
        return out

    @staticmethod
    def segments(out: dict, b: int):
        """(ph_idx_seq, ph_time_int, intervals) of utterance b from the dict ``run`` returned."""
        for b0, b1, seg_off, v in out["chunks"]:
            if b0 <= b < b1:
                j = b - b0
                o, k = int(seg_off[j]), int(v["n_seg"][j])
                return v["ph_idx_seq"][o:o + k], v["ph_time_int"][o:o + k], v["intervals"][o:o + k]
        raise IndexError(b)
