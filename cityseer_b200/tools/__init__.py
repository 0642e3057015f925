"""Shapely-free graph preparation helpers for the centrality path (inputs either side of the kernels)."""
from . import graphs, io, mock  # noqa: F401
