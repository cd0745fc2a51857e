"""the batched inference driver (GPU) and the TextGrid / confidence
writers (host).  Reference: infer.py:52-70, forced_alignment.py:154-186, tools/export_tool.py."""
import pathlib

import numpy as np
import pytest

from hubertfa_b200 import export


def _records(tmp):
    return [
        [tmp / "a" / "u1.wav", 2.0, np.float32(0.87654321), np.array(["SP", "a", "b", "SP"]),
         np.array([[0.0, 0.25], [0.25, 0.75], [0.75, 1.5], [1.5, 2.0]]),
         np.array(["SP", "ab", "SP"]), np.array([[0.0, 0.25], [0.25, 1.5], [1.5, 2.0]])],
        [tmp / "a" / "u2.wav", 1.0, np.float32(0.5), np.array(["k", "AP"]), np.array([[0.1, 0.4], [0.6, 1.0]]),
         np.array(["k", "AP"]), np.array([[0.1, 0.4], [0.6, 1.0]])],
    ]


def test_textgrid_layout_and_roundtrip(tmp_path):
    recs = _records(tmp_path)
    export.Exporter(recs, [], out_path=tmp_path / "out").export(["textgrid"])
    tg = tmp_path / "out" / "TextGrid" / "u2.TextGrid"
    text = tg.read_text()
    assert text.startswith('File type = "ooTextFile"\nObject class = "TextGrid"\n\nxmin = 0.0\nxmax = 1.0\n'
                           'tiers? <exists>\nsize = 2\nitem []:\n\titem [1]:\n\t\tclass = "IntervalTier"\n'
                           '\t\tname = "words"\n')
    tiers = export.read_textgrid(tg)
    assert list(tiers) == ["words", "phones"]                      # export_tool.py:31-32 order
    # gaps (leading 0..0.1 and 0.4..0.6) become empty-text intervals; the tier then tiles [0, xmax]
    assert tiers["phones"] == [(0.0, 0.1, ""), (0.1, 0.4, "k"), (0.4, 0.6, ""), (0.6, 1.0, "AP")]
    t1 = export.read_textgrid(tmp_path / "out" / "TextGrid" / "u1.TextGrid")
    assert [m for _, _, m in t1["words"]] == ["SP", "ab", "SP"]
    assert [(a, b) for a, b, _ in t1["phones"]] == [(0.0, 0.25), (0.25, 0.75), (0.75, 1.5), (1.5, 2.0)]


def test_textgrid_rejects_empty_and_overlapping_intervals():
    with pytest.raises(ValueError):     # textgrid.Interval: minTime >= maxTime
        export.textgrid_text(["a"], [[0.5, 0.5]], ["a"], [[0.5, 0.5]])
    with pytest.raises(ValueError):     # IntervalTier.addInterval: overlap
        export.textgrid_text(["a", "b"], [[0.0, 0.6], [0.5, 1.0]], ["a"], [[0.0, 1.0]])


def test_default_output_folder_and_confidence_csv_matches_pandas(tmp_path):
    pd = pytest.importorskip("pandas")
    recs = _records(tmp_path)
    export.Exporter(recs, ["some error"]).export(["textgrid", "confidence"])
    assert (tmp_path / "a" / "TextGrid" / "u1.TextGrid").is_file()            # export_tool.py:38-39
    got = (tmp_path / "a" / "confidence" / "confidence.csv").read_text()
    want = pd.DataFrame({"name": ["u1", "u2"], "confidence": [r[2] for r in recs]}).to_csv(index=False)
    assert got == want                                                          # export_tool.py:75-81


@pytest.mark.gpu
def test_batched_predictor_equals_per_utterance_decode():
    """trainer.predict (one decode per utterance) vs BatchedPredictor (one decode_batch per bucket):
    identical records, through post_processing, for a stand-in network head."""
    import torch
    from hubertfa_b200 import synth
    from hubertfa_b200.alignment_decoder import AlignmentDecoder
    from hubertfa_b200.infer import BatchedPredictor, split_head
    from hubertfa_b200.post_processing import post_processing

    V = 63
    T = np.array([310, 500, 257, 90, 700, 64], dtype=np.int32)
    S = np.array([31, 40, 65, 9, 150, 12], dtype=np.int32)
    vocab, items = synth.make_batch(T, S, V, seed=77, style="dictionary", planted=True)
    dev = torch.device("cuda")
    torch.manual_seed(0)
    head = torch.nn.Linear(16, V + 2).to(dev)                 # stand-in for UNet + linear head
    feats = [torch.randn(1, int(t), 16, device=dev) for t in T]
    # make the stand-in head peaked along each utterance's planted path so that paths are stable
    extra = [torch.cat([it["edge"][0][:, None], torch.zeros(int(t), 1), it["frame"][0]], dim=1).to(dev)[None]
             for it, t in zip(items, T)]
    forward = lambda k: head(feats[k]) * 0.05 + extra[k]
    wav_len = [float(t) * synth.FRAME_SECONDS - 0.004 for t in T]
    dataset = [(pathlib.Path(f"/x/u{k}.wav"), wav_len[k], k, it["ph_seq"], it["word_seq"], it["ph_idx_to_word_idx"])
               for k, it in enumerate(items)]
    dec = AlignmentDecoder(vocab, synth.MELSPEC_50FPS)
    pred = BatchedPredictor(forward, dec, bucket_utts=4)
    got = pred.predict(dataset)
    assert pred.n_buckets == 2 and len(got) == len(items)
    want = []
    with torch.no_grad():
        for wav_path, wl, k, ph, wd, p2w in dataset:            # forced_alignment.py:154-186
            frame, edge = split_head(forward(k))
            want.append((wav_path, wl, *[None] * 0) + tuple())
            r = dec.decode(frame, edge, None, wl, ph, wd, p2w)
            want[-1] = (wav_path, wl, r[4], r[0], r[1], r[2], r[3])
    for g, w in zip(got, want):
        assert g[0] == w[0] and g[1] == w[1] and g[2] == w[2]
        assert list(g[3]) == list(w[3]) and list(g[5]) == list(w[5])
        assert np.array_equal(g[4], w[4]) and np.array_equal(g[6], w[6])
    # ... and against the ORACLE: the reference's own arithmetic (CPU torch softmax + the NumPy restatement of
    # decode(), wav-length trimming and SP filter / word merge included) on the logits the network produced
    from oracle import hfa_oracle_np as onp
    with torch.no_grad():
        for g, (wav_path, wl, k, ph, wd, p2w) in zip(got, dataset):
            frame, edge = split_head(forward(k))
            exp = onp.decode(vocab, synth.MELSPEC_50FPS, frame.float().cpu(), edge.float().cpu(), None, wl, ph, wd, p2w)
            assert list(g[3]) == list(exp[0]) and list(g[5]) == list(exp[2]), wav_path
            np.testing.assert_allclose(g[4], exp[1], rtol=0, atol=1e-7)
            np.testing.assert_allclose(g[6], exp[3], rtol=0, atol=1e-7)
            np.testing.assert_allclose(g[2], exp[4], rtol=1e-4)
    # one unknown phoneme must cost one utterance, not the bucket
    bad = list(dataset[1])
    bad[3] = ["SP", "no-such-phoneme", "SP"]
    pred2 = BatchedPredictor(forward, dec, bucket_utts=4)
    got2 = pred2.predict([dataset[0], tuple(bad), dataset[2]])
    assert len(got2) == 2 and len(pred2.error_log) == 1 and "KeyError" in pred2.error_log[0][1]
    res_g, log_g = post_processing([list(x) for x in got])
    res_w, log_w = post_processing([list(x) for x in want])
    assert not log_g and not log_w
    for a, b in zip(res_g, res_w):
        assert list(a[3]) == list(b[3]) and np.array_equal(a[4], b[4]) and np.array_equal(a[6], b[6])
