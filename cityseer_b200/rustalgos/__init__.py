"""Drop-in for the part of ``cityseer.rustalgos`` that the centrality hot path uses.

Threshold pairing follows /root/reference/rust/src/common.rs:99-270 with every intermediate evaluated in IEEE f32,
rounding half away from zero like Rust's ``f32::round``.
"""
from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np

from . import centrality, graph  # noqa: F401  (submodules, as in cityseer.rustalgos)

_F32 = np.float32
MIN_THRESH_WT = _F32(0.01831563888873418)
WALKING_SPEED = _F32(1.33333)


def _f32(x) -> np.float32:
    return _F32(x)


def _ln_f32(x: np.float32) -> np.float32:
    # f32 natural log: correctly rounded from the f64 evaluation of the f32 argument
    return _F32(math.log(float(x)))


def _round_f32(x: np.float32) -> np.float32:
    """Rust ``f32::round``: nearest integer, ties away from zero."""
    v = float(x)
    if math.isnan(v) or math.isinf(v):
        return _F32(v)
    r = math.floor(abs(v) + 0.5)
    return _F32(math.copysign(r, v))


def _as_u32_list(name: str, vals) -> list[int]:
    if vals is None or isinstance(vals, (str, bytes)) or not isinstance(vals, (Sequence, np.ndarray)):
        raise TypeError(f"argument '{name}': expected a sequence of unsigned integers")
    out = []
    for v in vals:
        if isinstance(v, (bool, np.bool_)) or not isinstance(v, (int, np.integer)):
            raise TypeError(f"argument '{name}': '{type(v).__name__}' object cannot be interpreted as an integer")
        if v < 0 or v > 0xFFFFFFFF:
            raise OverflowError(f"argument '{name}': out of range integral type conversion attempted")
        out.append(int(v))
    return out


def _as_f32_list(name: str, vals) -> list[np.float32]:
    if vals is None or isinstance(vals, (str, bytes)) or not isinstance(vals, (Sequence, np.ndarray)):
        raise TypeError(f"argument '{name}': expected a sequence of floats")
    out = []
    for v in vals:
        if isinstance(v, (str, bytes)) or not isinstance(v, (int, float, np.integer, np.floating)):
            raise TypeError(f"argument '{name}': must be real number, not {type(v).__name__}")
        out.append(_F32(v))
    return out


def distances_from_betas(betas, min_threshold_wt=None) -> list[int]:
    """common.rs:99-133"""
    b = _as_f32_list("betas", betas)
    if len(b) == 0:
        raise ValueError("Input 'betas' cannot be empty.")
    mtw = MIN_THRESH_WT if min_threshold_wt is None else _F32(min_threshold_wt)
    if any(b[i + 1] >= b[i] for i in range(len(b) - 1)):
        raise ValueError("Betas must be unique and sorted in strictly decreasing order.")
    out = []
    ln = _ln_f32(mtw)
    for beta in b:
        if beta <= 0.0:
            raise ValueError("Beta values must be greater than zero.")
        d = _round_f32(ln / -beta)
        if d <= 0.0:
            raise ValueError("Derived distance must be positive. Check beta values.")
        out.append(int(d))
    return out


def betas_from_distances(distances, min_threshold_wt=None) -> list[float]:
    """common.rs:135-167"""
    d = _as_u32_list("distances", distances)
    if len(d) == 0:
        raise ValueError("Input 'distances' cannot be empty.")
    mtw = MIN_THRESH_WT if min_threshold_wt is None else _F32(min_threshold_wt)
    if any(d[i + 1] <= d[i] for i in range(len(d) - 1)):
        raise ValueError("Distances must be unique and sorted in strictly increasing order.")
    out = []
    neg_ln = -_ln_f32(mtw)
    for dist in d:
        if dist == 0:
            raise ValueError("Distances must be positive integers.")
        beta = neg_ln / _F32(dist)
        out.append(float(_round_f32(beta * _F32(100000.0)) / _F32(100000.0)))
    return out


def distances_from_seconds(seconds, speed_m_s) -> list[int]:
    """common.rs:169-203"""
    s = _as_u32_list("seconds", seconds)
    speed = _F32(speed_m_s)
    if len(s) == 0:
        raise ValueError("Input 'seconds' cannot be empty.")
    if speed <= 0.0:
        raise ValueError("Speed must be positive.")
    if any(s[i + 1] <= s[i] for i in range(len(s) - 1)):
        raise ValueError("Times must be unique and sorted in strictly increasing order.")
    out = []
    for t in s:
        if t == 0:
            raise ValueError("Time values must be positive integers.")
        d = _round_f32(_F32(t) * speed)
        if d <= 0.0:
            raise ValueError("Derived distance must be positive. Check time and speed values.")
        out.append(int(d))
    return out


def seconds_from_distances(distances, speed_m_s) -> list[int]:
    """common.rs:205-237"""
    d = _as_u32_list("distances", distances)
    speed = _F32(speed_m_s)
    if len(d) == 0:
        raise ValueError("Input 'distances' cannot be empty.")
    if speed <= 0.0:
        raise ValueError("Speed must be positive.")
    if any(d[i + 1] <= d[i] for i in range(len(d) - 1)):
        raise ValueError("Distances must be unique and sorted in strictly increasing order.")
    out = []
    for dist in d:
        if dist == 0:
            raise ValueError("Distances must be positive integers.")
        t = _round_f32(_F32(dist) / speed)
        if t <= 0.0:
            raise ValueError("Derived time must be positive. Check distance and speed values.")
        out.append(int(t))
    return out


def pair_distances_betas_time(speed_m_s, distances=None, betas=None, minutes=None, min_threshold_wt=None):
    """common.rs:239-270 — exactly one of distances / betas / minutes."""
    mtw = MIN_THRESH_WT if min_threshold_wt is None else _F32(min_threshold_wt)
    if distances is not None and betas is None and minutes is None:
        d = _as_u32_list("distances", distances)
        b = betas_from_distances(d, mtw)
        s = seconds_from_distances(d, speed_m_s)
        return d, b, s
    if distances is None and betas is not None and minutes is None:
        b = [float(x) for x in _as_f32_list("betas", betas)]
        d = distances_from_betas(b, mtw)
        s = seconds_from_distances(d, speed_m_s)
        return d, b, s
    if distances is None and betas is None and minutes is not None:
        m = _as_f32_list("minutes", minutes)
        s = [int(_round_f32(x * _F32(60.0))) for x in m]
        d = distances_from_seconds(s, speed_m_s)
        b = betas_from_distances(d, mtw)
        return d, b, s
    raise ValueError("Please provide exactly one of the following arguments: 'distances', 'betas', or 'minutes'.")
