"""hubertfa_b200 -- B200-native (sm_100a) forced-alignment decoder, drop-in for HubertFA's
``tools/alignment_decoder.py``.  See DESIGN.md and include/hfa_align.h."""
from ._lib import HfaError, LIB_PATH  # noqa: F401

__all__ = ["AlignmentDecoder", "BatchAlignment", "AlignPlan", "HfaError", "LIB_PATH"]


def __getattr__(name):
    # torch is imported lazily so that `import hubertfa_b200` stays cheap for pure-host users
    if name in ("AlignmentDecoder", "BatchAlignment"):
        from . import alignment_decoder
        return getattr(alignment_decoder, name)
    if name == "AlignPlan":
        from .ops import AlignPlan
        return AlignPlan
    raise AttributeError(name)
