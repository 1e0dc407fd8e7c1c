"""clairs_to_b200 -- B200-native (sm_100a) drop-in for the ClairS-TO per-candidate hot path:
pileup tensor encoder -> AFF (CvT) / NEG (BiGRU) forward -> posterior combine.

Everything numeric runs in ``libcto_b200.so`` (hand-written CUDA behind the C ABI of
``include/clairs_to_b200.h``); Python is host orchestration only.  See DESIGN.md.
"""

__version__ = "0.1.0"

from .pileup_format import PileupStream, N_POS, N_CH  # noqa: F401


def load_engine():
    """Import the engine lazily (needs torch + the built CUDA library)."""
    from .engine import Engine
    return Engine
