"""Batched inference driver: the caller of the decoder, widened to whole buckets of utterances.

Reference: ``infer.py:52-70`` + ``LitForcedAlignmentTask.predict_step``
(networks/task/forced_alignment.py:154-186).  There ``trainer.predict`` walks the dataset one
utterance per step, and every step ends in ``decoder.decode`` -- four device->host copies and a
host-side DP per utterance.  Here the network still runs utterance by utterance (it is out of scope
and stays as it is), but its head outputs STAY ON THE DEVICE and a whole bucket of them is aligned
by ONE ``decode_batch`` call (one collation, four kernel launches, one result download).  The
records that come out have the layout ``predict_step`` returns (:178-186), so
``post_processing`` (tools/post_processing.py:68) and the exporter take them unchanged.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

import numpy as np
import torch

from .alignment_decoder import AlignmentDecoder

__all__ = ["BatchedPredictor", "split_head"]


def split_head(logits: torch.Tensor):
    """The three views ``LitForcedAlignmentTask.forward`` cuts out of the head output
    (forced_alignment.py:288-291): frame logits = columns 2.., edge logit = column 0.  The ctc
    stream (column 1 and 3..) is only used by validation and is not materialised here.
    logits: [1, T, V+2] or [T, V+2].  The views are strided -- the emission kernel reads them in
    place."""
    if logits.dim() == 3:
        return logits[:, :, 2:], logits[:, :, 0]
    return logits[:, 2:], logits[:, 0]


class BatchedPredictor:
    """``predictor.predict(dataset)`` == ``trainer.predict(model, dataset)`` of the reference.

    forward: callable(item) -> head logits ``[1, T, V+2]`` (device tensor), or the tuple
             ``(ph_frame_logits, ph_edge_logits[, ctc_logits])`` the reference's ``forward``
             returns.  It owns wav loading / units encoding / the network (all out of scope here).
    dataset: iterable of ``(wav_path, wav_length, features, ph_seq, word_seq, ph_idx_to_word_idx)``;
             ``features`` is whatever ``forward`` needs, ``wav_length`` in seconds or None.
    bucket_cells / bucket_utts: a bucket is flushed to ``decode_batch`` when its DP cells
             (sum T*S) or its utterance count reach these bounds -- they bound the device memory
             held by un-aligned logits and by the alignment workspace (~8.5 B per cell).
    """

    def __init__(self, forward: Callable, decoder: AlignmentDecoder, bucket_cells: float = 3e8,
                 bucket_utts: int = 4096):
        self.forward = forward
        self.decoder = decoder
        self.bucket_cells = float(bucket_cells)
        self.bucket_utts = int(bucket_utts)
        self.n_buckets = 0
        self.error_log = []          # (wav_path, message) of utterances that could not be aligned

    def _flush(self, bucket, out):
        if not bucket:
            return
        res = self.decoder.decode_batch(
            [b[1] for b in bucket], [b[2] for b in bucket], [b[0][3] for b in bucket],
            [b[0][4] for b in bucket], [b[0][5] for b in bucket], wav_lengths=[b[0][1] for b in bucket])
        self.n_buckets += 1
        for j, (item, _, _) in enumerate(bucket):
            try:
                ph_seq, ph_intervals, word_seq, word_intervals, confidence = res[j]
            except Exception as e:                   # one bad utterance (empty after trimming, ...) must not
                self.error_log.append((item[0], f"{type(e).__name__}: {e}"))      # take the bucket down:
                continue                             # logged like post_processing does (post_processing.py:82-104)
            out.append((item[0], item[1], confidence, ph_seq, ph_intervals, word_seq, word_intervals))
        bucket.clear()

    @torch.no_grad()
    def predict(self, dataset: Iterable[Sequence]) -> list:
        out, bucket, cells = [], [], 0.0
        for item in dataset:
            wav_path, wav_length, features, ph_seq, word_seq, ph_idx_to_word_idx = item
            try:
                self.decoder._ids_of(ph_seq)         # unknown phoneme / bad id: this utterance only
            except (KeyError, IndexError) as e:
                self.error_log.append((wav_path, f"{type(e).__name__}: {e}"))
                continue
            y = self.forward(features)
            if isinstance(y, (tuple, list)):
                frame, edge = y[0], y[1]
            else:
                frame, edge = split_head(y)
            if word_seq is None:                     # alignment_decoder.py:41-43
                word_seq, ph_idx_to_word_idx = ph_seq, np.arange(len(ph_seq))
            T = frame.shape[-2]
            bucket.append(((wav_path, wav_length, None, ph_seq, word_seq, ph_idx_to_word_idx), frame, edge))
            cells += float(T) * len(ph_seq)
            if cells >= self.bucket_cells or len(bucket) >= self.bucket_utts:
                self._flush(bucket, out)
                cells = 0.0
        self._flush(bucket, out)
        return out
