"""TextGrid / confidence export of the aligned intervals -- SURVEY.md 8(f) rank 4.

Reference: ``tools/export_tool.py`` (``Exporter.save_textgrids`` :13-45, ``save_confidence_fn``
:47-81, ``export`` :83-92), fed by ``post_processing`` (infer.py:63-70).  The reference builds
``textgrid.TextGrid`` objects and ``pandas`` frames; neither wire format needs those packages, so
both are written directly here (the ``textgrid`` package is not installed in this image):

* TextGrid: Praat "long" text format exactly as ``textgrid.TextGrid.write`` lays it out -- tier
  ``words`` then ``phones`` (:18-19,31-32), gaps between intervals filled with empty-text
  intervals, numbers printed with ``'{0}'.format(float)``.
* confidence.csv: ``name,confidence`` rows per wav folder (:75-81), as ``DataFrame.to_csv(index=False)``
  prints them.

PARITY UNPINNED for the TextGrid text: the ``textgrid`` package is absent from the build container, so
the layout follows its writer as documented/remembered and the Praat specification; the tests
round-trip it through ``read_textgrid`` below and check the gap-filling and ordering rules.
"""
from __future__ import annotations

import pathlib
import re
from typing import Iterable, Sequence

__all__ = ["Exporter", "write_textgrid", "read_textgrid", "textgrid_text"]


def _fill_gaps(intervals, t_min, t_max):
    """``IntervalTier._fillInTheGaps``: empty-text intervals between neighbours and up to t_max."""
    out, prev = [], t_min
    for lo, hi, mark in intervals:
        if prev < lo:
            out.append((prev, lo, ""))
        out.append((lo, hi, mark))
        prev = hi
    if t_max is not None and prev < t_max:
        out.append((prev, t_max, ""))
    return out


def _tier(name, marks, intervals):
    rows = []
    for mark, (lo, hi) in zip(marks, intervals):
        lo, hi = float(lo), float(hi)
        if lo >= hi:                               # textgrid.Interval raises the same way
            raise ValueError((lo, hi))
        rows.append((lo, hi, str(mark)))
    rows.sort(key=lambda r: r[0])
    for a, b in zip(rows, rows[1:]):               # IntervalTier.addInterval refuses overlaps
        if b[0] < a[1]:
            raise ValueError(f"overlapping intervals in tier {name!r}: {a} / {b}")
    return name, rows


def textgrid_text(word_seq, word_intervals, ph_seq, ph_intervals) -> str:
    """The file ``Exporter.save_textgrids`` writes for one utterance (export_tool.py:17-45)."""
    tiers = [_tier("words", word_seq, word_intervals), _tier("phones", ph_seq, ph_intervals)]
    t_max = max((rows[-1][1] for _, rows in tiers if rows), default=0.0)
    fmt = "{0}".format
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "xmin = " + fmt(0.0),
           "xmax = " + fmt(t_max), "tiers? <exists>", "size = " + fmt(len(tiers)), "item []:"]
    for i, (name, rows) in enumerate(tiers, 1):
        filled = _fill_gaps(rows, 0.0, None)
        out += [f"\titem [{i}]:", '\t\tclass = "IntervalTier"', f'\t\tname = "{name}"',
                "\t\txmin = " + fmt(0.0), "\t\txmax = " + fmt(t_max),
                f"\t\tintervals: size = {len(filled)}"]
        for j, (lo, hi, mark) in enumerate(filled, 1):
            out += [f"\t\t\tintervals [{j}]:", "\t\t\t\txmin = " + fmt(lo), "\t\t\t\txmax = " + fmt(hi),
                    '\t\t\t\ttext = "{0}"'.format(mark.replace('"', '""'))]
    return "\n".join(out) + "\n"


def write_textgrid(path, word_seq, word_intervals, ph_seq, ph_intervals) -> None:
    path = pathlib.Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(textgrid_text(word_seq, word_intervals, ph_seq, ph_intervals), encoding="utf-8")


_NUM = re.compile(r"^\s*(xmin|xmax)\s*=\s*([-+0-9.eE]+)\s*$")
_TXT = re.compile(r'^\s*text\s*=\s*"(.*)"\s*$')
_NAME = re.compile(r'^\s*name\s*=\s*"(.*)"\s*$')


def read_textgrid(path_or_text) -> dict:
    """Minimal reader of the long format written above: {tier name: [(xmin, xmax, text), ...]}.
    (What ``tools/label.py:63-71`` needs back from a TextGrid; used by the round-trip tests.)"""
    text = path_or_text if "\n" in str(path_or_text) else pathlib.Path(path_or_text).read_text(encoding="utf-8")
    tiers, cur, lo, hi = {}, None, None, None
    in_interval = False
    for line in text.splitlines():
        m = _NAME.match(line)
        if m:
            cur = tiers.setdefault(m.group(1), [])
            in_interval = False
            continue
        if re.match(r"^\s*intervals \[\d+\]:", line):
            in_interval = True
            continue
        m = _NUM.match(line)
        if m and in_interval:
            if m.group(1) == "xmin":
                lo = float(m.group(2))
            else:
                hi = float(m.group(2))
            continue
        m = _TXT.match(line)
        if m and in_interval and cur is not None:
            cur.append((lo, hi, m.group(1).replace('""', '"')))
            in_interval = False
    return tiers


class Exporter:
    """Same constructor and ``export(out_formats)`` as the reference class (export_tool.py:7-92).
    predictions: the records ``post_processing`` returns -- ``[wav_path, wav_length, confidence,
    ph_seq, ph_intervals, word_seq, word_intervals]``."""

    def __init__(self, predictions: Iterable[Sequence], log, out_path=None):
        self.predictions = predictions
        self.log = log
        self.out_path = pathlib.Path(out_path) if out_path else None

    def save_textgrids(self):
        print("Saving TextGrids...")
        for wav_path, _, _, ph_seq, ph_intervals, word_seq, word_intervals in self.predictions:
            wav_path = pathlib.Path(wav_path)
            name = wav_path.with_suffix(".TextGrid").name
            base = self.out_path if self.out_path is not None else wav_path.parent   # :36-39
            write_textgrid(base / "TextGrid" / name, word_seq, word_intervals, ph_seq, ph_intervals)

    def save_confidence_fn(self):
        print("saving confidence...")
        folders = {}
        for wav_path, _, confidence, *_ in self.predictions:
            wav_path = pathlib.Path(wav_path)
            folders.setdefault(wav_path.parent, []).append((wav_path.with_suffix("").name, confidence))
        for folder, rows in folders.items():                                       # :75-81
            path = folder / "confidence"
            path.mkdir(parents=True, exist_ok=True)
            # a list of np.float32 scalars (the decoder's total_confidence) becomes a float32 column, which
            # pandas prints with the shortest float32 representation == str(np.float32)
            lines = ["name,confidence"] + [f"{_csv(n)},{str(c) if hasattr(c, 'dtype') else repr(float(c))}"
                                           for n, c in rows]
            (path / "confidence.csv").write_text("\n".join(lines) + "\n", encoding="utf-8")

    def export(self, out_formats):
        self.save_textgrids()
        if "confidence" in out_formats:
            self.save_confidence_fn()
        if self.log:
            print("error:")
            for line in self.log:
                print(line)


def _csv(field: str) -> str:
    field = str(field)
    if any(c in field for c in ',"\n\r'):
        return '"' + field.replace('"', '""') + '"'
    return field
