"""cityseer_b200 — B200-native localized network centrality behind cityseer's NetworkStructure API.

Only the hot path named in BASELINE.json is implemented: ``centrality_shortest``,
``centrality_simplest`` and ``segment_centrality`` run as hand-written sm_100a CUDA kernels reached through the
``extern "C"`` library declared in ``include/cityseer_b200.h``.  There is no CPU fallback: if the CUDA library is
missing or no GPU is present, the compute entry points raise.
"""
from . import config, rustalgos  # noqa: F401

__version__ = "0.1.0"
