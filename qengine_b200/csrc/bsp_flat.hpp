// The trace tree as the device sees it, built on the host from the on-disk BSP lumps (MakeTnodes, trace.c:47-88;
// breadth-first here), plus the box cells of the leaves for the cell proofs (walkers.cuh).  Host only.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/qrad_cuda.h"

namespace qrad {

struct FlatBsp {
  std::vector<int4> nodes;        // x = float bits of dist, y = plane type, z = child 0, w = child 1 (walkers.cuh: Tree)
  std::vector<float4> normals;    // xyz = plane normal, w = dist
  std::vector<int32_t> leaf_contents, leaf_cluster;
  std::vector<float4> cell_lo, cell_hi;   // per leaf; cell_lo.w = 1: solid leaf whose ancestors are all axial planes
  int depth = 0;                  // nodes on the longest root -> leaf path
  float extent = 0;               // largest |dist| of an axial node plane
};

// throws std::runtime_error on malformed input
void flatten_bsp(const qrad_bsp_view *b, FlatBsp &out);

// how deep inside a solid cell a segment must pass to count as proven blocked: ON_EPSILON + slack for the rounding of
// the clip points (<= ~2 ulp of the largest coordinate per split)
inline float proof_margin(float extent) {
  const float e = extent > 1.f ? extent : 1.f;
  const float slack = 300.f * e * 1.2e-7f;
  return 0.1f + (slack > 0.15f ? slack : 0.15f);
}

}  // namespace qrad
