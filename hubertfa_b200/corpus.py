"""Corpus-scale alignment over the GPUs of one box: utterances sharded by cost, streamed in bounded chunks,
results gathered ON THE HOST without a collective.

The DP is independent per utterance (the reference decodes one utterance per call,
networks/task/forced_alignment.py:154-186; infer.py:60-61 is a batch-size-1 loop), so multi-GPU is pure
data parallelism: one process per GPU, no NCCL on the data path.  Every rank

  1. takes its shard of the corpus (``sharding.shard_by_cost``: longest-processing-time-first over T*S),
  2. cuts it into chunks of bounded workspace (``sharding.chunk_by_bytes``), collates every chunk ONCE
     (``hfa_plan_create`` + ``hfa_plan_upload`` into the chunk's own workspace) and
  3. per pass launches ``hfa_align_batch`` per chunk, alternating between two streams so that the tail of
     one chunk overlaps the start of the next, and copies each chunk's compact result blob (status, n_seg,
     ph_idx_seq, ph_time_int, intervals, confidence: ~24 bytes per phoneme) device -> host.

The host side of the gather is a POSIX shared-memory segment that every rank maps and registers with CUDA
(``cudaHostRegister``): each rank's D2H copies land directly in its slice, at offsets every rank derives from
the corpus description alone, so after a barrier rank 0 simply reads all results in corpus order -- nothing is
pickled, nothing crosses a socket, NVLink is not involved (there is nothing for it to do: results are ~1 %
of the logits' bytes).
"""
from __future__ import annotations

from multiprocessing import shared_memory

import numpy as np
import torch

from . import _lib, ops, sharding


class SharedHostBuffer:
    """A named shared-memory segment mapped by every rank of the box and page-locked for CUDA."""

    def __init__(self, name: str, nbytes: int, create: bool):
        self.nbytes = max(int(nbytes), 4096)
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=self.nbytes)
        if not create:
            # Python < 3.13 registers attached segments with the resource tracker too, which would unlink
            # the owner's segment when this process exits: only the creating rank may unlink it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((self.nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.nbytes, 0)
            if int(getattr(rc, "value", rc)) != 0:
                raise _lib.HfaError(f"cudaHostRegister of the shared result segment failed ({rc})")
            self._registered = True
        self.owner = create

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = self.array = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


class CorpusPlan:
    """Everything that follows from the corpus description alone -- identical on every rank."""

    def __init__(self, T, S, ids_list, vocab_size: int, frame_length: float, world: int, max_cells: int):
        self.T = np.ascontiguousarray(T, dtype=np.int32)
        self.S = np.ascontiguousarray(S, dtype=np.int32)
        self.ids_list = ids_list
        self.vocab_size, self.frame_length, self.world = int(vocab_size), float(frame_length), int(world)
        self.n_utt = int(self.T.size)
        self.cells = int((self.T.astype(np.int64) * self.S).sum())
        self.frames = int(self.T.sum(dtype=np.int64))
        self.shards = sharding.shard_by_cost(self.T, self.S, world)
        # chunks[r] = list of corpus-index arrays; inside a chunk longest utterance first (collation order)
        self.chunks = []
        for shard in self.shards:
            # equal-sized chunks (a small remainder chunk would leave one of the two streams idle): as few as the
            # workspace bound allows, balanced by cost like the shards themselves; inside a chunk longest first
            cells = int((self.T[shard].astype(np.int64) * self.S[shard]).sum())
            n_chunks = max(1, -(-cells // max(int(max_cells), 1)))
            parts = sharding.shard_by_cost(self.T[shard], self.S[shard], n_chunks)
            chunks = []
            for part in parts:
                idx = shard[part]
                if len(idx):
                    chunks.append(idx[np.argsort(-self.T[idx].astype(np.int64), kind="stable")])
            self.chunks.append(chunks)
        # result blob of a chunk: the HfaResultLayout of its plan; its size follows from (n_utt, sum S) alone
        self.blob_off, off = [], 0
        for r in range(world):
            offs = []
            for idx in self.chunks[r]:
                offs.append(off)
                off += self.result_bytes(len(idx), int(self.S[idx].sum()))
            self.blob_off.append(offs)
        self.total_result_bytes = off

    @staticmethod
    def result_bytes(n_utt: int, n_states: int) -> int:
        """Mirror of the result-blob layout in hfa_plan_create (five per-utterance int32/float arrays, two
        per-state int32 arrays, one per-state double[2]; every region padded to 16 bytes)."""
        a16 = lambda b: (b + 15) // 16 * 16
        return max(5 * a16(4 * n_utt) + 2 * a16(4 * n_states) + a16(16 * n_states), 16) + 240   # slack: alignment

    def cells_of_rank(self, r: int) -> int:
        idx = self.shards[r]
        return int((self.T[idx].astype(np.int64) * self.S[idx]).sum())


class CorpusAligner:
    """One rank's share of a CorpusPlan: persistent per-chunk plans / workspaces, ``run()`` = one pass."""

    def __init__(self, cp: CorpusPlan, rank: int, device, head: torch.Tensor, row_off: np.ndarray,
                 host: SharedHostBuffer, frame_col: int = 2, edge_col: int = 0, n_streams: int = 2):
        """head: [sum T, W] logits of the WHOLE corpus (or at least of this rank's utterances at their corpus
        rows) on ``device``; row_off[b] = first row of utterance b."""
        self.cp, self.rank, self.dev, self.host = cp, rank, torch.device(device), host
        self.dtype = ops.TORCH_TO_DTYPE[head.dtype]
        self.chunks = []
        W, esz, base = head.shape[1], head.element_size(), head.data_ptr()
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(max(int(n_streams), 1))]
            for k, idx in enumerate(cp.chunks[rank]):
                ids = np.concatenate([cp.ids_list[i] for i in idx]) if len(idx) else np.zeros(0, np.int32)
                plan = ops.AlignPlan(cp.T[idx], cp.S[idx], ids, cp.vocab_size, cp.frame_length)
                if plan.result_bytes > cp.result_bytes(len(idx), int(cp.S[idx].sum())):
                    raise _lib.HfaError("result blob larger than the corpus plan reserved")
                ws, res = plan.new_workspace(self.dev), plan.new_result(self.dev)
                plan.upload(ws)
                rows = row_off[idx].astype(np.int64)
                n = len(idx)
                plan.set_inputs(ws, base + (rows * W + frame_col) * esz, np.full(n, W), np.ones(n),
                                base + (rows * W + edge_col) * esz, np.full(n, W))
                o = cp.blob_off[rank][k]
                self.chunks.append(dict(idx=idx, plan=plan, ws=ws, res=res,
                                        host=host.tensor[o:o + plan.result_bytes]))
            torch.cuda.synchronize(self.dev)
        self.launches_per_pass = None

    def run(self) -> None:
        """One pass over this rank's shard: kernels + D2H of every chunk's results into the shared host
        segment, enqueued on two alternating streams behind the current stream; returns when enqueued --
        ``join()`` makes the current stream wait for it."""
        cur = torch.cuda.current_stream(self.dev)
        n0 = ops.launch_count()
        for st in self.streams:
            st.wait_stream(cur)
        for k, c in enumerate(self.chunks):
            st = self.streams[k % len(self.streams)]
            with torch.cuda.stream(st):
                ops.align_batch(c["ws"], c["plan"].handle, self.dtype, c["res"], None)
                c["host"].copy_(c["res"][:c["plan"].result_bytes], non_blocking=True)
        self.launches_per_pass = ops.launch_count() - n0

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)

    def close(self):
        for c in self.chunks:
            c["plan"].close()
        self.chunks = []


def read_results(cp: CorpusPlan, host: SharedHostBuffer, utterances=None):
    """Rank 0 after the barrier: per-utterance (status, ph_idx_seq, ph_time_int, intervals, total_conf,
    final_score) straight out of the shared segment, for ``utterances`` (corpus indices; default all)."""
    want = None if utterances is None else set(int(u) for u in utterances)
    out = {}
    a16 = lambda b: (b + 15) // 16 * 16
    for r in range(cp.world):
        for k, idx in enumerate(cp.chunks[r]):
            if want is not None and not want.intersection(int(i) for i in idx):
                continue
            n, ns = len(idx), int(cp.S[idx].sum())
            blob = host.array[cp.blob_off[r][k]:]
            o = 0
            fields = {}
            for name, dt, cnt in (("status", np.int32, n), ("n_seg", np.int32, n), ("end_state", np.int32, n),
                                  ("final_score", np.float32, n), ("total_conf", np.float32, n),
                                  ("ph_idx_seq", np.int32, ns), ("ph_time_int", np.int32, ns),
                                  ("intervals", np.float64, 2 * ns)):
                nb = cnt * np.dtype(dt).itemsize
                fields[name] = blob[o:o + nb].view(dt)
                o += a16(nb)
            seg_off = np.concatenate([[0], np.cumsum(cp.S[idx].astype(np.int64))])
            for j, u in enumerate(idx):
                if want is not None and int(u) not in want:
                    continue
                s0, kk = int(seg_off[j]), int(fields["n_seg"][j])
                out[int(u)] = dict(status=int(fields["status"][j]), ph_idx_seq=fields["ph_idx_seq"][s0:s0 + kk].copy(),
                                   ph_time_int=fields["ph_time_int"][s0:s0 + kk].copy(),
                                   intervals=fields["intervals"][2 * s0:2 * (s0 + kk)].reshape(-1, 2).copy(),
                                   total_conf=float(fields["total_conf"][j]), final_score=float(fields["final_score"][j]))
    return out


def all_status_ok(cp: CorpusPlan, host: SharedHostBuffer) -> bool:
    """True when every utterance of every rank's chunks reports status 0 (reads only the status words)."""
    for r in range(cp.world):
        for k, idx in enumerate(cp.chunks[r]):
            o = cp.blob_off[r][k]
            if host.array[o:o + 4 * len(idx)].view(np.int32).any():
                return False
    return True
