// centrality_simplest kernels (placeholder until the shortest path is validated on hardware).
#pragma once
#include "cs_common.cuh"
