"""Drop-in for the centrality functions of ``cityseer.metrics.networks``
(/root/reference/pysrc/cityseer/metrics/networks.py:101-275, :464-635, :638-755, thin wrappers :763-888).

Same arguments, same threshold resolution, same sampled / full split, same result columns
(``cc_{measure}_{d}``, ``_ang`` suffix for simplest, ``cc_seg_*`` for segment, ``hillier = density² / farness``); the
frame is any ``pandas.DataFrame`` indexed by node key (a GeoDataFrame works the same — geopandas is not required)."""
from __future__ import annotations

import logging
from functools import partial

import numpy as np
import pandas as pd

from .. import config, rustalgos, sampling

logger = logging.getLogger(__name__)
MIN_THRESH_WT = config.MIN_THRESH_WT
SPEED_M_S = config.SPEED_M_S


def _require_dual_for_angular(network_structure, context: str) -> None:
    # networks.py:90-98
    if not network_structure.is_dual:
        raise ValueError(
            f"{context} requires a dual graph for angular analysis. Convert the graph with "
            "cityseer.tools.graphs.nx_to_dual(...) before ingesting it into NetworkStructure."
        )


def _split_sampled(resolved, sample: bool, epsilon):
    eps = epsilon if epsilon is not None else sampling.HOEFFDING_EPSILON
    full, sampled = [], []
    if not sample:
        return sorted(resolved), sampled
    logger.warning("Sampling is experimental: API and behaviour may change in future releases.")
    for d in sorted(resolved):
        p = sampling.compute_distance_p(d, epsilon=eps)
        if p >= 1.0:
            full.append(d)
        else:
            sampled.append((d, p))
    return full, sampled


def _run_split(method, network_structure, full, sampled, common: dict, random_seed):
    node_count = network_structure.street_node_count()
    results = {}
    if full:
        label = ", ".join(f"{d}m" for d in full)
        logger.info(f"  Full: {label}")
        fn = partial(method, distances=full, random_seed=random_seed, **common)
        res = config.wrap_progress(total=node_count, rust_struct=network_structure, partial_func=fn, desc=f"centrality full: {label}")
        for d in full:
            results[d] = res
    for d, p in sampled:
        logger.info(f"  Sampled {d}m: p={p:.0%}")
        fn = partial(method, distances=[d], sample_probability=p, random_seed=random_seed, **common)
        results[d] = config.wrap_progress(total=node_count, rust_struct=network_structure, partial_func=fn,
                                          desc=f"centrality p={p:.0%}: {d}m")  # fmt: skip
    return results


def _write(nodes_gdf: pd.DataFrame, temp_data: dict, node_keys_py) -> pd.DataFrame:
    temp_df = pd.DataFrame(temp_data, index=node_keys_py)
    gdf_idx = nodes_gdf.index.intersection(node_keys_py)
    for col in temp_df.columns:
        if col not in nodes_gdf.columns:
            nodes_gdf[col] = np.nan
    nodes_gdf.loc[gdf_idx, temp_df.columns] = temp_df.loc[gdf_idx, temp_df.columns]
    return nodes_gdf


def node_centrality_shortest(network_structure, nodes_gdf, distances=None, betas=None, minutes=None,
                             compute_closeness: bool = True, compute_betweenness: bool = True,
                             min_threshold_wt: float = MIN_THRESH_WT, speed_m_s: float = SPEED_M_S, tolerance=None,
                             random_seed=None, sample: bool = False, epsilon=None):  # fmt: skip
    """networks.py:101-275"""
    logger.info("Computing node centrality (shortest).")
    resolved, _b, _s = rustalgos.pair_distances_betas_time(speed_m_s, distances, betas, minutes, min_threshold_wt=min_threshold_wt)
    full, sampled = _split_sampled(resolved, sample, epsilon)
    common = dict(compute_closeness=compute_closeness, compute_betweenness=compute_betweenness,
                  min_threshold_wt=min_threshold_wt, speed_m_s=speed_m_s, tolerance=tolerance)  # fmt: skip
    results = _run_split(network_structure.centrality_shortest, network_structure, full, sampled, common, random_seed)
    if not results:
        return nodes_gdf
    temp_data: dict[str, object] = {}
    if compute_closeness:
        for key, attr in [("beta", "node_beta"), ("cycles", "node_cycles"), ("density", "node_density"),
                          ("farness", "node_farness"), ("harmonic", "node_harmonic")]:  # fmt: skip
            for d, res in results.items():
                temp_data[config.prep_gdf_key(key, d)] = getattr(res, attr)[d]
        for d, res in results.items():
            with np.errstate(divide="ignore", invalid="ignore"):
                temp_data[config.prep_gdf_key("hillier", d)] = res.node_density[d] ** 2 / res.node_farness[d]
    if compute_betweenness:
        for key, attr in [("betweenness", "node_betweenness"), ("betweenness_beta", "node_betweenness_beta")]:
            for d, res in results.items():
                temp_data[config.prep_gdf_key(key, d)] = getattr(res, attr)[d]
    return _write(nodes_gdf, temp_data, next(iter(results.values())).node_keys_py)


def node_centrality_simplest(network_structure, nodes_gdf, distances=None, betas=None, minutes=None,
                             compute_closeness: bool = True, compute_betweenness: bool = True,
                             min_threshold_wt: float = MIN_THRESH_WT, speed_m_s: float = SPEED_M_S,
                             angular_scaling_unit: float = 90, farness_scaling_offset: float = 1, tolerance=None,
                             random_seed=None, sample: bool = False, epsilon=None):  # fmt: skip
    """networks.py:464-635 (wrapper default ``angular_scaling_unit=90``; the native default is 180)."""
    _require_dual_for_angular(network_structure, "node_centrality_simplest")
    logger.info("Computing node centrality (simplest).")
    resolved, _b, _s = rustalgos.pair_distances_betas_time(speed_m_s, distances, betas, minutes, min_threshold_wt=min_threshold_wt)
    full, sampled = _split_sampled(resolved, sample, epsilon)
    common = dict(compute_closeness=compute_closeness, compute_betweenness=compute_betweenness,
                  min_threshold_wt=min_threshold_wt, speed_m_s=speed_m_s, tolerance=tolerance,
                  angular_scaling_unit=angular_scaling_unit, farness_scaling_offset=farness_scaling_offset)  # fmt: skip
    results = _run_split(network_structure.centrality_simplest, network_structure, full, sampled, common, random_seed)
    if not results:
        return nodes_gdf
    temp_data: dict[str, object] = {}
    if compute_closeness:
        for d, res in results.items():
            temp_data[config.prep_gdf_key("density", d, angular=True)] = res.node_density[d]
            temp_data[config.prep_gdf_key("harmonic", d, angular=True)] = res.node_harmonic[d]
            temp_data[config.prep_gdf_key("farness", d, angular=True)] = res.node_farness[d]
            with np.errstate(divide="ignore", invalid="ignore"):
                temp_data[config.prep_gdf_key("hillier", d, angular=True)] = res.node_density[d] ** 2 / res.node_farness[d]
    if compute_betweenness:
        for d, res in results.items():
            temp_data[config.prep_gdf_key("betweenness", d, angular=True)] = res.node_betweenness[d]
    return _write(nodes_gdf, temp_data, next(iter(results.values())).node_keys_py)


def segment_centrality(network_structure, nodes_gdf, distances=None, betas=None, minutes=None,
                       compute_closeness: bool = True, compute_betweenness: bool = True,
                       min_threshold_wt: float = MIN_THRESH_WT, speed_m_s: float = SPEED_M_S):  # fmt: skip
    """networks.py:638-755"""
    logger.info("Computing shortest path segment centrality.")
    fn = partial(network_structure.segment_centrality, distances=distances, betas=betas, minutes=minutes,
                 compute_closeness=compute_closeness, compute_betweenness=compute_betweenness,
                 min_threshold_wt=min_threshold_wt, speed_m_s=speed_m_s)  # fmt: skip
    result = config.wrap_progress(total=network_structure.street_node_count(), rust_struct=network_structure, partial_func=fn)
    resolved, b_, s_ = rustalgos.pair_distances_betas_time(speed_m_s, distances, betas, minutes, min_threshold_wt=min_threshold_wt)
    config.log_thresholds(resolved, b_, s_)
    temp_data = {}
    if compute_closeness is True:
        for key, attr in [("seg_density", "segment_density"), ("seg_harmonic", "segment_harmonic"), ("seg_beta", "segment_beta")]:
            for d in resolved:
                temp_data[config.prep_gdf_key(key, d)] = getattr(result, attr)[d]
    if compute_betweenness is True:
        for d in resolved:
            temp_data[config.prep_gdf_key("seg_betweenness", d)] = result.segment_betweenness[d]
    return _write(nodes_gdf, temp_data, result.node_keys_py)


def betweenness_od(network_structure, nodes_gdf, od_matrix, distances=None, betas=None, minutes=None,
                   min_threshold_wt: float = MIN_THRESH_WT, speed_m_s: float = SPEED_M_S, tolerance=None):  # fmt: skip
    """OD-weighted betweenness (networks.py:383-461): only origins with outbound trips are searched, every shortest
    path contribution is scaled by its OD weight; writes ``cc_betweenness_{d}`` / ``cc_betweenness_beta_{d}``."""
    logger.info("Computing OD-weighted betweenness centrality.")
    fn = partial(network_structure.betweenness_od_shortest, od_matrix=od_matrix, distances=distances, betas=betas,
                 minutes=minutes, min_threshold_wt=min_threshold_wt, speed_m_s=speed_m_s, tolerance=tolerance)  # fmt: skip
    result = config.wrap_progress(total=network_structure.street_node_count(), rust_struct=network_structure, partial_func=fn)
    resolved, _b, _s = rustalgos.pair_distances_betas_time(speed_m_s, distances, betas, minutes, min_threshold_wt)
    temp_data = {}
    for measure_key, attr_key in (("betweenness", "node_betweenness"), ("betweenness_beta", "node_betweenness_beta")):
        for d in resolved:
            temp_data[config.prep_gdf_key(measure_key, d)] = getattr(result, attr_key)[d]
    return _write(nodes_gdf, temp_data, result.node_keys_py)


# ---- closeness-only / betweenness-only conveniences (networks.py:763-888)
def closeness_shortest(network_structure, nodes_gdf, **kw):
    return node_centrality_shortest(network_structure, nodes_gdf, compute_closeness=True, compute_betweenness=False, **kw)


def betweenness_shortest(network_structure, nodes_gdf, **kw):
    return node_centrality_shortest(network_structure, nodes_gdf, compute_closeness=False, compute_betweenness=True, **kw)


def closeness_simplest(network_structure, nodes_gdf, **kw):
    return node_centrality_simplest(network_structure, nodes_gdf, compute_closeness=True, compute_betweenness=False, **kw)


def betweenness_simplest(network_structure, nodes_gdf, **kw):
    return node_centrality_simplest(network_structure, nodes_gdf, compute_closeness=False, compute_betweenness=True, **kw)
