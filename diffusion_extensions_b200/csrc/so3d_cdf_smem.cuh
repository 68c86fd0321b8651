// so3d_cdf_smem.cuh -- the CDF row of one timestep staged in shared memory (shared-t fast paths of the sampler and
// of the reverse step, distributions.py:33-51 with a scalar eps): grid locations, the row's trapezoid CDF and a
// float-format guide (so3d_math.cuh: sg_bucket) built by the CTA, so that an inverse-CDF lookup costs ~1 probe instead of a
// 10-step search.
#pragma once

#include "so3d_math.cuh"

namespace so3d {

// tab layout: [loc kGrid][trap kGrid][guide kGuideStride u16 = kGuideStride/2 floats (+1)][5 scalars]
constexpr int kTabLoc = 0, kTabTrap = kGrid, kTabGuide = 2 * kGrid, kTabScal = 2 * kGrid + kGuideStride / 2 + 1;
constexpr int kTabCdfFloats = kTabScal + 8;

// loc (and, for a shared row, the CDF row and its guide) staged by the whole CTA
__device__ __forceinline__ void stage_cdf(float* tab, const float* __restrict__ cdf_row, const float* __restrict__ loc) {
  for (int k = threadIdx.x; k < kCdf; k += blockDim.x) {
    tab[kTabLoc + k] = loc[k];
    if (cdf_row) tab[kTabTrap + k] = cdf_row[k];
  }
  __syncthreads();
  if (cdf_row) {
    uint16_t* guide = reinterpret_cast<uint16_t*>(tab + kTabGuide);
    for (int k = threadIdx.x; k <= kSgBuckets; k += blockDim.x)
      guide[k] = (uint16_t)cdf_count_le(tab + kTabTrap, sg_edge(k), 0, kCdf);
  }
}
__device__ __forceinline__ float shared_row_angle(const float* tab, float u) {
  return igso3_angle_from_uniform_guided(tab + kTabTrap, tab + kTabLoc, reinterpret_cast<const uint16_t*>(tab + kTabGuide), u);
}

}  // namespace so3d
