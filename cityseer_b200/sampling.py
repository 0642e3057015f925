"""Distance-based sampling schedule (Hoeffding / Eppstein-Wang) used by the ``sample=True`` mode of the wrappers.

Formulas follow /root/reference/pysrc/cityseer/sampling.py:32-104: k = ln(2r/δ) / (2ε²), p = min(1, k/r), with the
canonical grid model r = π d² / s²."""
from __future__ import annotations

import math

HOEFFDING_EPSILON: float = 0.06
HOEFFDING_DELTA: float = 0.1
GRID_SPACING: float = 175.0


def compute_hoeffding_p(mean_reachability: float, epsilon: float = HOEFFDING_EPSILON, delta: float = HOEFFDING_DELTA) -> float:
    vals = (mean_reachability, epsilon, delta)
    if any(not math.isfinite(v) for v in vals) or mean_reachability <= 0 or epsilon <= 0 or delta <= 0 or delta >= 1:
        return 1.0
    k = math.log(2 * mean_reachability / delta) / (2 * epsilon**2)
    return min(1.0, k / mean_reachability)


def compute_distance_p(distance: float, epsilon: float = HOEFFDING_EPSILON, delta: float = HOEFFDING_DELTA,
                       grid_spacing: float = GRID_SPACING) -> float:  # fmt: skip
    if distance <= 0 or grid_spacing <= 0:
        return 1.0
    r = math.pi * distance**2 / grid_spacing**2
    return compute_hoeffding_p(r, epsilon=epsilon, delta=delta)
