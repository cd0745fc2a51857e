"""One small batch through every forward route (strips incl. exchange, bands, one warp per utterance), checked
against the oracle -- the program compute-sanitizer runs (tools/gpu/r2_sanitize.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gpu_util import run_core_gpu, check_core_against_oracle, synth_core_inputs
shapes = [(70, 20, "dictionary"), (90, 45, "dictionary"), (130, 100, "alternate"), (40, 7, "nosp"), (3, 7, "alternate")]
ins = [synth_core_inputs(T, S, 63, 4000 + i, style, planted=bool(i % 2)) for i, (T, S, style) in enumerate(shapes)]
for mode, kern in (("2", "skew"), ("2", "band"), ("0", "skew")):
    os.environ["HFA_LATENCY_MODE"], os.environ["HFA_LAT_KERNEL"] = mode, kern
    out = run_core_gpu([x["ids"] for x in ins], [x["prob_log"] for x in ins], [x["el"] for x in ins],
                       [x["ne"] for x in ins], [x["p"] for x in ins], 0.02)
    for x, g in zip(ins, out):
        check_core_against_oracle(x["ids"], x["prob_log"], x["el"], x["ne"], g, x["p"], 0.02)
    print("route", mode, kern, "ok", flush=True)
