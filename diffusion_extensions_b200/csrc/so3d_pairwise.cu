// so3d_pairwise.cu -- all-pairs kernel sums over two sets of rotations: the arithmetic of the reference's
// MMD two-sample statistic (util.py:254-285) with rmat_gaussian_kernel (util.py:128-134, exp(-rmat_dist),
// rmat_dist = |log(A^T B)|_F = sqrt(2) theta, util.py:315-322) or rmat_cosine_kernel (util.py:136-150, cos theta).
//
// The reference materialises chunk x chunk x 3 x 3 products (chunksize 4000 at bingham_test.py:29, 20 000^2 pairs
// per sum) and runs log_rmat (with its eigh fallback) on every pair.  Here nothing is materialised:
//   * every rotation becomes a unit quaternion once per tile visit (40 instructions, amortised over 256 pairs);
//   * a CTA owns a 256 x 256 tile of pairs: each thread keeps ONE row quaternion in registers and walks the 256
//     column quaternions of the tile in shared memory (LDS.128 broadcast: one wavefront per warp and pair);
//   * per pair: relative quaternion a* (x) b (16 FMA), sin(theta/2) = |vector part| computed directly and
//     cos(theta/2) = |scalar part| (no 1 - d^2 cancellation: theta is accurate to ~2e-7 rad from 0 to pi, like
//     the reference's atan2 formulation and unlike an acos of the trace), theta/2 = atan2 by a first-quadrant
//     polynomial, exp through MUFU.EX2: ~41 FP32-pipe instructions + 3 MUFU -> FP32-issue bound;
//   * X-X and Y-Y sums visit only the lower-triangular tile pairs (off-diagonal tiles weigh 2);
//   * per-thread fp32 partials over one tile column sweep (<= 256 terms <= 1), accumulated across tiles in
//     double, reduced per CTA into a caller-provided scratch array and summed in a fixed order by a second
//     tiny kernel: the result is deterministic for a given (shard, nshards, scratch length).
// Multi-GPU: the flat list of tile pairs is dealt round-robin to `nshards` shards; each rank computes its share and
// the three partial sums are all-reduced by the caller (parallel.py) -- the one real reduction of this path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/so3d.h"
#include "so3d_common.cuh"
#include "so3d_math.cuh"

using namespace so3d;

namespace {

constexpr int kPT = 256;  // rows (and columns) per tile == threads per CTA

__device__ __forceinline__ float4 load_quat(const float* __restrict__ R, int64_t row, int64_t n) {
  if (row >= n) return make_float4(1.f, 0.f, 0.f, 0.f);
  Mat3 m;
#pragma unroll
  for (int k = 0; k < 9; ++k) m.m[k] = __ldg(R + row * 9 + k);
  float q[4];
  rmat_to_quat(m, q);
  return make_float4(q[0], q[1], q[2], q[3]);
}

// value of the kernel for the pair of unit quaternions (real part in .x); arithmetic in so3d_math.cuh
template <int kKernel>
__device__ __forceinline__ float pair_value(const float4 a, const float4 b) {
  return so3_pair_kernel<kKernel == SO3D_PAIR_GAUSSIAN>(Quat{a.x, a.y, a.z, a.w}, Quat{b.x, b.y, b.z, b.w});
}

struct PairArgs {
  const float* X;
  const float* Y;
  int64_t nx, ny;
  int64_t tx, ty;          // tiles per set
  int64_t pxx, pyy, pxy;   // tile pairs per section: X-X (lower triangle), Y-Y (lower triangle), X-Y (all)
  int64_t shard, nshards;
  double* ws;
};

// q in [0, T(T+1)/2) -> (i, j), j <= i, row-major lower triangle
__device__ __forceinline__ void tri_decode(int64_t q, int64_t* i_out, int64_t* j_out) {
  int64_t i = (int64_t)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
  while (i * (i + 1) / 2 > q) --i;
  while ((i + 1) * (i + 2) / 2 <= q) ++i;
  *i_out = i;
  *j_out = q - i * (i + 1) / 2;
}

template <int kKernel>
__global__ void __launch_bounds__(kPT) pair_sums_kernel(const PairArgs a) {
  __shared__ float4 s_q[kPT];
  __shared__ double s_red[3][kPT / 32];
  const int tid = threadIdx.x;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;  // X-X, Y-Y, X-Y
  const int64_t total = a.pxx + a.pyy + a.pxy;
  for (int64_t p = a.shard + a.nshards * (int64_t)blockIdx.x; p < total; p += a.nshards * (int64_t)gridDim.x) {
    int sec;
    int64_t bi, bj;
    if (p < a.pxx) {
      sec = 0;
      tri_decode(p, &bi, &bj);
    } else if (p < a.pxx + a.pyy) {
      sec = 1;
      tri_decode(p - a.pxx, &bi, &bj);
    } else {
      sec = 2;
      const int64_t q = p - a.pxx - a.pyy;
      bi = q / a.ty;
      bj = q - bi * a.ty;
    }
    const float* A = sec == 1 ? a.Y : a.X;
    const float* B = sec == 0 ? a.X : a.Y;
    const int64_t na = sec == 1 ? a.ny : a.nx, nb = sec == 0 ? a.nx : a.ny;
    const int64_t row = bi * kPT + tid;
    const float4 qa = load_quat(A, row, na);
    __syncthreads();  // the previous tile's column sweep is over
    s_q[tid] = load_quat(B, bj * kPT + tid, nb);
    __syncthreads();
    const int64_t left = nb - bj * kPT;
    const int cols = (int)(left < kPT ? left : kPT);
    float part = 0.f;
    if (cols == kPT) {
#pragma unroll 8
      for (int j = 0; j < kPT; ++j) part += pair_value<kKernel>(qa, s_q[j]);
    } else {
      for (int j = 0; j < cols; ++j) part += pair_value<kKernel>(qa, s_q[j]);
    }
    const double w = row < na ? (double)((sec < 2 && bi != bj) ? 2.0f * part : part) : 0.0;
    acc0 += sec == 0 ? w : 0.0;
    acc1 += sec == 1 ? w : 0.0;
    acc2 += sec == 2 ? w : 0.0;
  }
  // CTA reduction (fixed order): lanes by shuffle, warps through shared memory
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    double v = s == 0 ? acc0 : (s == 1 ? acc1 : acc2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_red[s][tid >> 5] = v;
  }
  __syncthreads();
  if (tid < 3) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kPT / 32; ++w) v += s_red[tid][w];
    a.ws[3 * (int64_t)blockIdx.x + tid] = v;
  }
}

// out3[s] = sum over the CTAs' partials, one warp per section, fixed order
__global__ void __launch_bounds__(96) pair_sums_finish(const double* __restrict__ ws, int ctas, double* __restrict__ out3) {
  const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int c = lane; c < ctas; c += 32) v += ws[3 * (int64_t)c + s];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out3[s] = v;
}

template <int kKernel>
int launch_pair_sums(const PairArgs& a, int64_t ws_len, double* out3, void* stream) {
  static int resident = 0;
  if (resident == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_sums_kernel<kKernel>, kPT, 0) != cudaSuccess || occ < 1) occ = 1;
    resident = occ;
  }
  const int64_t total = a.pxx + a.pyy + a.pxy;
  const int64_t mine = total > a.shard ? (total - a.shard + a.nshards - 1) / a.nshards : 0;
  int64_t grid = (int64_t)so3d_host::sm_count() * resident;
  if (grid > mine) grid = mine;
  if (grid > ws_len / 3) grid = ws_len / 3;
  if (grid < 1) grid = 1;
  pair_sums_kernel<kKernel><<<(int)grid, kPT, 0, (cudaStream_t)stream>>>(a);
  if (int rc = so3d_host::check_launch("so3d_pair_kernel_sums_f32")) return rc;
  pair_sums_finish<<<1, 96, 0, (cudaStream_t)stream>>>(a.ws, (int)grid, out3);
  return so3d_host::check_launch("so3d_pair_kernel_sums_f32 (finish)");
}

}  // namespace

extern "C" int so3d_pair_kernel_sums_f32(const float* X, int64_t nx, const float* Y, int64_t ny, int kernel, int64_t shard,
                                         int64_t nshards, double* ws, int64_t ws_len, double* out3, void* stream) {
  if (nx < 0 || ny < 0) return so3d_host::fail(SO3D_EINVAL, "so3d_pair_kernel_sums_f32: negative size");
  if (!ws || ws_len < 3 || !out3) return so3d_host::fail(SO3D_EINVAL, "so3d_pair_kernel_sums_f32: ws (>= 3 doubles) and out3 are required");
  if ((nx > 0 && !X) || (ny > 0 && !Y)) return so3d_host::fail(SO3D_EINVAL, "so3d_pair_kernel_sums_f32: null pointer");
  if (nshards < 1 || shard < 0 || shard >= nshards) return so3d_host::fail(SO3D_EINVAL, "so3d_pair_kernel_sums_f32: need 0 <= shard < nshards");
  if (kernel != SO3D_PAIR_GAUSSIAN && kernel != SO3D_PAIR_COSINE)
    return so3d_host::fail(SO3D_EINVAL, "so3d_pair_kernel_sums_f32: kernel must be SO3D_PAIR_GAUSSIAN or SO3D_PAIR_COSINE");
  PairArgs a;
  a.X = X; a.Y = Y; a.nx = nx; a.ny = ny;
  a.tx = (nx + kPT - 1) / kPT; a.ty = (ny + kPT - 1) / kPT;
  a.pxx = a.tx * (a.tx + 1) / 2; a.pyy = a.ty * (a.ty + 1) / 2; a.pxy = a.tx * a.ty;
  a.shard = shard; a.nshards = nshards; a.ws = ws;
  return kernel == SO3D_PAIR_GAUSSIAN ? launch_pair_sums<SO3D_PAIR_GAUSSIAN>(a, ws_len, out3, stream)
                                      : launch_pair_sums<SO3D_PAIR_COSINE>(a, ws_len, out3, stream);
}
