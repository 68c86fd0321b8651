// so3d_kernels.cu -- sm_100a kernels + C ABI (include/so3d.h) for the SO(3) diffusion hot path.
//
// Execution model (B200: 148 SMs, 64 warps/SM, 227 KB smem/CTA, HBM3e ~6.5 TB/s measured):
//   * one rotation per thread, all 3x3 math in registers, every op of a call fused in one kernel;
//   * rotations are 36-byte AoS records, so a warp reading "its" rows directly would touch each
//     128-byte line 9 times.  Instead a CTA moves a whole tile (kTile rows, contiguous in memory)
//     between HBM and shared memory with coalesced 16-byte vector accesses and each thread reads its
//     row from shared memory at a stride of 9 (or 3) words -- odd, hence bank-conflict free;
//   * grids are persistent: min(#tiles, 148 * CTAs/SM) CTAs striding over the tiles;
//   * HBM-bound kernels use streaming loads/stores (ld.global.cs / st.global.cs): nothing is re-read;
//   * the series evaluator is bound by FP32 operand bandwidth / MUFU / issue at once (7 FP32 + 1 MUFU.EX2 +
//     1 uniform constant load per term, see so3d_math.cuh).
// No library calls, no tensor cores (nothing here is GEMM shaped).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/so3d.h"
#include "so3d_math.cuh"
#include "so3d_tma.cuh"

using namespace so3d;

namespace {

constexpr int kTile = 256;  // rows per tile == threads per CTA

thread_local char g_err[256] = "";

int fail(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}

int check_launch(const char* name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      g_sm_count = 148;
  }
  return g_sm_count;
}

inline int grid_for(int64_t n, int ctas_per_sm) {
  const int64_t tiles = (n + kTile - 1) / kTile;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// tile movers: contiguous block of rows*W floats between global and shared memory
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void tile_load(float* __restrict__ sm, const float* __restrict__ g, int rows, bool vec) {
  const int nwords = rows * W;
  if (vec) {
    const int nvec = nwords >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(sm);
    for (int i = threadIdx.x; i < nvec; i += kTile) s4[i] = __ldcs(g4 + i);
    for (int i = (nvec << 2) + threadIdx.x; i < nwords; i += kTile) sm[i] = __ldcs(g + i);
  } else {
    for (int i = threadIdx.x; i < nwords; i += kTile) sm[i] = __ldcs(g + i);
  }
}

template <int W>
__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* __restrict__ sm, int rows, bool vec) {
  const int nwords = rows * W;
  if (vec) {
    const int nvec = nwords >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* s4 = reinterpret_cast<const float4*>(sm);
    for (int i = threadIdx.x; i < nvec; i += kTile) __stcs(g4 + i, s4[i]);
    for (int i = (nvec << 2) + threadIdx.x; i < nwords; i += kTile) __stcs(g + i, sm[i]);
  } else {
    for (int i = threadIdx.x; i < nwords; i += kTile) __stcs(g + i, sm[i]);
  }
}

__device__ __forceinline__ Mat3 sm_mat(const float* sm, int r) {
  Mat3 m;
#pragma unroll
  for (int k = 0; k < 9; ++k) m.m[k] = sm[r * 9 + k];
  return m;
}
__device__ __forceinline__ void sm_put_mat(float* sm, int r, const Mat3& m) {
#pragma unroll
  for (int k = 0; k < 9; ++k) sm[r * 9 + k] = m.m[k];
}
__device__ __forceinline__ Vec3 sm_vec(const float* sm, int r) { return Vec3{sm[r * 3], sm[r * 3 + 1], sm[r * 3 + 2]}; }
__device__ __forceinline__ void sm_put_vec(float* sm, int r, Vec3 v) {
  sm[r * 3] = v.x;
  sm[r * 3 + 1] = v.y;
  sm[r * 3 + 2] = v.z;
}

// Generic row-wise kernel.  Op provides:
//   static constexpr int kIn9, kIn3, kOut9, kOut3   -- how many n x 9 / n x 3 arrays it reads / writes
//   const float* in9[kIn9], in3[kIn3]; float* out9[kOut9], out3[kOut3]  (an output may be NULL = skipped)
//   __device__ void row(int64_t i, const Mat3* a9, const Vec3* a3, Mat3* o9, Vec3* o3) const
// Per-row scalars (angles, eps, t, ...) are read/written directly by Op::row with coalesced accesses.
template <class Op>
__global__ void __launch_bounds__(kTile) rowwise_kernel(const Op op, const int64_t n, const unsigned vecmask) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  constexpr int kI9 = Op::kIn9, kI3 = Op::kIn3, kO9 = Op::kOut9, kO3 = Op::kOut3;
  float* s_i9 = smem;
  float* s_i3 = s_i9 + kI9 * kTile * 9;
  float* s_o9 = s_i3 + ((kI3 * kTile * 3 + 3) & ~3);
  float* s_o3 = s_o9 + kO9 * kTile * 9;
  const int64_t tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int rows = (int)((n - row0) < kTile ? (n - row0) : kTile);
    unsigned bit = 0;
#pragma unroll
    for (int a = 0; a < kI9; ++a, ++bit) tile_load<9>(s_i9 + a * kTile * 9, op.in9[a] + row0 * 9, rows, (vecmask >> bit) & 1u);
#pragma unroll
    for (int a = 0; a < kI3; ++a, ++bit) tile_load<3>(s_i3 + a * kTile * 3, op.in3[a] + row0 * 3, rows, (vecmask >> bit) & 1u);
    if (kI9 + kI3 > 0) __syncthreads();
    const int r = threadIdx.x;
    if (r < rows) {
      Mat3 a9[kI9 > 0 ? kI9 : 1];
      Vec3 a3[kI3 > 0 ? kI3 : 1];
      Mat3 o9[kO9 > 0 ? kO9 : 1];
      Vec3 o3[kO3 > 0 ? kO3 : 1];
#pragma unroll
      for (int a = 0; a < kI9; ++a) a9[a] = sm_mat(s_i9 + a * kTile * 9, r);
#pragma unroll
      for (int a = 0; a < kI3; ++a) a3[a] = sm_vec(s_i3 + a * kTile * 3, r);
      op.row(row0 + r, a9, a3, o9, o3);
#pragma unroll
      for (int a = 0; a < kO9; ++a) sm_put_mat(s_o9 + a * kTile * 9, r, o9[a]);
#pragma unroll
      for (int a = 0; a < kO3; ++a) sm_put_vec(s_o3 + a * kTile * 3, r, o3[a]);
    }
    if (kO9 + kO3 > 0) __syncthreads();
#pragma unroll
    for (int a = 0; a < kO9; ++a, ++bit)
      if (op.out9[a]) tile_store<9>(op.out9[a] + row0 * 9, s_o9 + a * kTile * 9, rows, (vecmask >> bit) & 1u);
#pragma unroll
    for (int a = 0; a < kO3; ++a, ++bit)
      if (op.out3[a]) tile_store<3>(op.out3[a] + row0 * 3, s_o3 + a * kTile * 3, rows, (vecmask >> bit) & 1u);
    // the next iteration's __syncthreads (after its loads) orders these smem reads before the next writes
    if (kI9 + kI3 == 0 && kO9 + kO3 > 0) __syncthreads();
  }
}

template <class Op>
constexpr size_t op_smem() {
  return sizeof(float) * (size_t)(Op::kIn9 * kTile * 9 + ((Op::kIn3 * kTile * 3 + 3) & ~3) + Op::kOut9 * kTile * 9 + Op::kOut3 * kTile * 3);
}

template <class Op>
int launch_rowwise(const Op& op, int64_t n, void* stream, const char* name, int ctas_per_sm = 8) {
  if (n < 0) return fail(SO3D_EINVAL, "negative n");
  if (n == 0) return 0;
  unsigned mask = 0, bit = 0;
  for (int a = 0; a < Op::kIn9; ++a, ++bit) {
    if (!op.in9[a]) return fail(SO3D_EINVAL, "null input pointer");
    mask |= (aligned16(op.in9[a]) ? 1u : 0u) << bit;
  }
  for (int a = 0; a < Op::kIn3; ++a, ++bit) {
    if (!op.in3[a]) return fail(SO3D_EINVAL, "null input pointer");
    mask |= (aligned16(op.in3[a]) ? 1u : 0u) << bit;
  }
  for (int a = 0; a < Op::kOut9; ++a, ++bit) mask |= (aligned16(op.out9[a]) ? 1u : 0u) << bit;
  for (int a = 0; a < Op::kOut3; ++a, ++bit) mask |= (aligned16(op.out3[a]) ? 1u : 0u) << bit;
  constexpr size_t smem = op_smem<Op>();
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    cudaFuncSetAttribute(rowwise_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  rowwise_kernel<Op><<<grid_for(n, ctas_per_sm), kTile, smem, (cudaStream_t)stream>>>(op, n, mask);
  return check_launch(name);
}

// dummy arrays for ops without a given kind of operand (zero-length arrays are not allowed)
#define SO3D_OP_ARRAYS(I9, I3, O9, O3)                                         \
  static constexpr int kIn9 = I9, kIn3 = I3, kOut9 = O9, kOut3 = O3;           \
  const float* in9[I9 > 0 ? I9 : 1];                                           \
  const float* in3[I3 > 0 ? I3 : 1];                                           \
  float* out9[O9 > 0 ? O9 : 1];                                                \
  float* out3[O3 > 0 ? O3 : 1];

// ------------------------------------------------------------------------------------------------
// L0 ops
// ------------------------------------------------------------------------------------------------
struct LogOp {  // util.py:164-192
  SO3D_OP_ARRAYS(1, 0, 1, 0)
  __device__ void row(int64_t, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const { o9[0] = hat(log_vec(a9[0])); }
};
struct LogVecOp {
  SO3D_OP_ARRAYS(1, 0, 0, 1)
  __device__ void row(int64_t, const Mat3* a9, const Vec3*, Mat3*, Vec3* o3) const { o3[0] = log_vec(a9[0]); }
};
struct RmatToAaOp {  // util.py:208-219
  SO3D_OP_ARRAYS(1, 0, 0, 1)
  float* angle;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3*, Vec3* o3) const {
    const AxisAngle a = axis_angle(a9[0]);
    o3[0] = a.axis;
    angle[i] = a.theta;
  }
};
struct AaToRmatOp {  // util.py:195-205
  SO3D_OP_ARRAYS(0, 1, 1, 0)
  const float* angle;
  __device__ void row(int64_t i, const Mat3*, const Vec3* a3, Mat3* o9, Vec3*) const { o9[0] = aa_to_rmat(a3[0], angle[i]); }
};
struct ExpVecOp {  // diffusion.py:294
  SO3D_OP_ARRAYS(0, 1, 1, 0)
  __device__ void row(int64_t, const Mat3*, const Vec3* a3, Mat3* o9, Vec3*) const { o9[0] = exp_vec(a3[0]); }
};
struct ScaleOp {  // util.py:349-361
  SO3D_OP_ARRAYS(1, 0, 1, 0)
  const float* s;
  int s_stride;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const { o9[0] = scale_rot(a9[0], s[i * s_stride]); }
};
struct RmatToQuatOp {
  SO3D_OP_ARRAYS(1, 0, 0, 0)
  float* q;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3*, Vec3*) const {
    float qq[4];
    rmat_to_quat(a9[0], qq);
    *reinterpret_cast<float4*>(q + 4 * i) = make_float4(qq[0], qq[1], qq[2], qq[3]);
  }
};
struct QuatToRmatOp {  // util.py:222-252
  SO3D_OP_ARRAYS(0, 0, 1, 0)
  const float* q;
  bool q_vec;
  __device__ void row(int64_t i, const Mat3*, const Vec3*, Mat3* o9, Vec3*) const {
    float4 v;
    if (q_vec) v = __ldcs(reinterpret_cast<const float4*>(q) + i);
    else v = make_float4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
    o9[0] = quat_to_rmat(v.x, v.y, v.z, v.w);
  }
};
template <bool TA, bool TB>
__device__ __forceinline__ Mat3 mul_op(const Mat3& a, const Mat3& b) {
  if (TA && TB) {  // A^T B^T = (B A)^T
    const Mat3 c = mul_nn(b, a);
    Mat3 t;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) t.m[3 * i + j] = c.m[3 * j + i];
    return t;
  }
  if (TA) return mul_tn(a, b);
  if (TB) return mul_nt(a, b);
  return mul_nn(a, b);
}
template <bool TA, bool TB>
struct ComposeOp {
  SO3D_OP_ARRAYS(2, 0, 1, 0)
  __device__ void row(int64_t, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const { o9[0] = mul_op<TA, TB>(a9[0], a9[1]); }
};
// one operand shared by all rows (e.g. mean @ R, distributions.py:50)
template <bool TA, bool TB, bool SharedIsA>
struct ComposeSharedOp {
  SO3D_OP_ARRAYS(1, 0, 1, 0)
  const float* shared;
  __device__ void row(int64_t, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const {
    Mat3 s;
#pragma unroll
    for (int k = 0; k < 9; ++k) s.m[k] = __ldg(shared + k);
    o9[0] = SharedIsA ? mul_op<TA, TB>(s, a9[0]) : mul_op<TA, TB>(a9[0], s);
  }
};
struct RmatDistOp {  // util.py:315-322
  SO3D_OP_ARRAYS(2, 0, 0, 0)
  float* out;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3*, Vec3*) const {
    out[i] = 1.41421356237f * axis_angle(mul_tn(a9[0], a9[1])).theta;
  }
};
struct LerpOp {  // util.py:325-338
  SO3D_OP_ARRAYS(2, 0, 1, 0)
  const float* w;
  int w_stride;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const {
    const AxisAngle a = axis_angle(mul_tn(a9[0], a9[1]));
    o9[0] = mul_nn(a9[0], rodrigues(a.axis, w[i * w_stride] * a.theta));
  }
};

// backward ops
struct LogBwdOp {
  SO3D_OP_ARRAYS(2, 0, 1, 0)
  __device__ void row(int64_t, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const { o9[0] = log_bwd(a9[0], a9[1]); }
};
struct AaToRmatBwdOp {
  SO3D_OP_ARRAYS(1, 1, 0, 1)
  const float* angle;
  float* g_angle;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3* a3, Mat3*, Vec3* o3) const {
    float ga;
    aa_to_rmat_bwd(a3[0], angle[i], a9[0], &o3[0], &ga);
    g_angle[i] = ga;
  }
};
struct ExpVecBwdOp {
  SO3D_OP_ARRAYS(1, 1, 0, 1)
  __device__ void row(int64_t, const Mat3* a9, const Vec3* a3, Mat3*, Vec3* o3) const {
    o3[0] = exp_vec_bwd(a3[0], exp_vec(a3[0]), a9[0]);
  }
};
struct ScaleBwdOp {
  // out = exp(hat(s * logvec(R))):  g_s = logvec . Jr^T u,  g_logvec = s Jr^T u, then through the log
  SO3D_OP_ARRAYS(2, 0, 1, 0)
  const float* s;
  int s_stride;
  float* g_s;  // per row (caller reduces for a shared scalar)
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const {
    const float sc = s[i * s_stride];
    const Vec3 lv = log_vec(a9[0]);
    const Vec3 w{sc * lv.x, sc * lv.y, sc * lv.z};
    const Vec3 gw = exp_vec_bwd(w, exp_vec(w), a9[1]);
    if (g_s) g_s[i] = fmaf(lv.x, gw.x, fmaf(lv.y, gw.y, lv.z * gw.z));
    const Vec3 gl{0.5f * sc * gw.x, 0.5f * sc * gw.y, 0.5f * sc * gw.z};  // GL = hat(g_logvec)/2
    o9[0] = log_bwd(a9[0], hat(gl));
  }
};

// ------------------------------------------------------------------------------------------------
// L1: IGSO(3)
// ------------------------------------------------------------------------------------------------
struct LogpScoreOp {  // distributions.py:74-77 + score, fused with the axis-angle extraction
  SO3D_OP_ARRAYS(1, 0, 0, 1)
  const float* eps;
  int eps_stride;
  float* logp;
  float* dlogf;
  int mode, L;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3*, Vec3* o3) const {
    const AxisAngle a = axis_angle(a9[0]);
    float lf, g;
    igso3_logf_g(a.theta, eps[i * eps_stride], mode, L, &lf, &g);
    logp[i] = lf;
    if (dlogf) dlogf[i] = g;
    o3[0] = Vec3{g * a.axis.x, g * a.axis.y, g * a.axis.z};
  }
};
struct LogpBwdOp {
  SO3D_OP_ARRAYS(1, 0, 1, 0)
  const float* dlogf;
  const float* gout;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const {
    const Mat3& r = a9[0];
    const float vx = r.m[7] - r.m[5], vy = r.m[2] - r.m[6], vz = r.m[3] - r.m[1];
    const float s = 0.5f * sqrtf(fmaf(vx, vx, fmaf(vy, vy, vz * vz)));
    const float c = 0.5f * (r.m[0] + r.m[4] + r.m[8] - 1.0f);
    const float gg = gout[i] * dlogf[i] / fmaf(s, s, c * c);
    // d theta / dR = [ c/(4 s) (R - R^T) - (s/2) I ] / (s^2 + c^2); at s -> 0 the skew part -> 0/0 * 0
    const float ka = s > 0.f ? gg * c / (4.0f * s) : 0.f;
    const float kd = -0.5f * gg * s;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q) o9[0].m[3 * p + q] = ka * (r.m[3 * p + q] - r.m[3 * q + p]) + (p == q ? kd : 0.f);
  }
};

__global__ void __launch_bounds__(256) density_kernel(const float* __restrict__ omega, const float* __restrict__ eps, int eps_stride,
                                                      float* __restrict__ f, int64_t n, int mode, int L) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float lf, g;
    igso3_logf_g(omega[i], eps[i * eps_stride], mode, L, &lf, &g);
    f[i] = expf(lf);
  }
}

// distributions.py:15-30.  One CTA per eps row: fp64 density at the 1000 grid points -> fp32, times
// the Haar weight (fp32), trapezoid increments (fp32), prefix sum accumulated in double and rounded
// to float per entry (what ATen's CPU cumsum does for float), normalised by the last entry.
__global__ void __launch_bounds__(256) cdf_table_kernel(const float* __restrict__ eps, const float* __restrict__ grid_loc,
                                                        const float* __restrict__ haar_w, float* __restrict__ trap_out, int quirks) {
  __shared__ float s_pdf[kGrid];
  __shared__ float s_inc[kCdf];
  __shared__ double s_part[8];
  const int row = blockIdx.x;
  const double e = (double)eps[row];
  for (int k = threadIdx.x; k < kGrid; k += blockDim.x) {
    const float loc = grid_loc[k];
    const float dens = (float)igso3_closed_f64((double)loc, e, quirks);  // :19-21 (.float() at :72)
    s_pdf[k] = (loc == 0.0f) ? 0.0f : dens * haar_w[k];                 // :21, :23
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kCdf; k += blockDim.x)
    s_inc[k] = (grid_loc[k + 1] - grid_loc[k]) * (s_pdf[k] + s_pdf[k + 1]) / 2.0f;  // :26-28
  __syncthreads();
  // 8 warps x 125 entries: serial double prefix inside a chunk (lane 0 of each warp), then offsets.
  // Summation order differs from a strictly serial loop only in double precision (error ~1e-16).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kChunk = (kCdf + 7) / 8;
  __shared__ double s_pre[kCdf];
  if (lane == 0) {
    double acc = 0.0;
    const int lo = warp * kChunk, hi = min(kCdf, lo + kChunk);
    for (int k = lo; k < hi; ++k) {
      acc += (double)s_inc[k];
      s_pre[k] = acc;
    }
    s_part[warp] = acc;
  }
  __syncthreads();
  double total = 0.0;
  for (int w = 0; w < 8; ++w) total += s_part[w];
  const float last = (float)total;
  for (int k = threadIdx.x; k < kCdf; k += blockDim.x) {
    double off = 0.0;
    const int w = k / kChunk;
    for (int j = 0; j < w; ++j) off += s_part[j];
    const float c = (float)(off + s_pre[k]);
    trap_out[(int64_t)row * kCdf + k] = c / last;  // :29
  }
}

// distributions.py:33-51.  kShared: all samples use one CDF row, staged in shared memory.
template <bool kShared>
__global__ void __launch_bounds__(kTile) sample_kernel(const float* __restrict__ cdf, const float* __restrict__ loc,
                                                       const int64_t* __restrict__ row_idx, int64_t row, int64_t rows,
                                                       const float* __restrict__ u_in, const float* __restrict__ axes_in,
                                                       uint64_t seed, uint64_t rng_offset, uint64_t row_offset,
                                                       const float* __restrict__ mean, int mean_stride, float* __restrict__ R,
                                                       float* __restrict__ angle_out, float* __restrict__ axis_out, int64_t n, bool r_vec) {
  extern __shared__ float4 smem4[];
  float* s_out = reinterpret_cast<float*>(smem4);  // kTile * 9
  float* s_loc = s_out + kTile * 9;                // 1000 (999 used)
  float* s_cdf = s_loc + 1000;                     // 1000 (999 used) when kShared
  for (int k = threadIdx.x; k < kCdf; k += kTile) {
    s_loc[k] = loc[k];
    if (kShared) s_cdf[k] = cdf[row * kCdf + k];
  }
  __syncthreads();
  const int64_t tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int rows_here = (int)((n - row0) < kTile ? (n - row0) : kTile);
    const int r = threadIdx.x;
    if (r < rows_here) {
      const int64_t i = row0 + r;
      Vec3 axis;
      float u;
      if (axes_in) {
        const float ax = axes_in[3 * i], ay = axes_in[3 * i + 1], az = axes_in[3 * i + 2];
        const float inv = 1.0f / sqrtf(fmaf(ax, ax, fmaf(ay, ay, az * az)));  // distributions.py:36
        axis = Vec3{ax * inv, ay * inv, az * inv};
      }
      if (!axes_in || !u_in) {
        const NoiseDraw d = draw_axis_u(seed, row_offset + (uint64_t)i, rng_offset);
        if (!axes_in) axis = d.axis;
        u = d.u;
      }
      if (u_in) u = u_in[i];
      const float* trap;
      if (kShared) {
        trap = s_cdf;
      } else {
        int64_t rr = row_idx[i];
        rr = rr < 0 ? 0 : (rr >= rows ? rows - 1 : rr);
        trap = cdf + rr * kCdf;
      }
      const float ang = igso3_angle_from_uniform(trap, s_loc, u);
      Mat3 out = rodrigues(axis, ang);
      if (mean) {
        Mat3 m;
        const float* mp = mean + (mean_stride ? 9 * i : 0);
#pragma unroll
        for (int k = 0; k < 9; ++k) m.m[k] = __ldg(mp + k);
        out = mul_nn(m, out);
      }
      sm_put_mat(s_out, r, out);
      if (angle_out) angle_out[i] = ang;
      if (axis_out) {
        axis_out[3 * i] = axis.x;
        axis_out[3 * i + 1] = axis.y;
        axis_out[3 * i + 2] = axis.z;
      }
    }
    __syncthreads();
    tile_store<9>(R + row0 * 9, s_out, rows_here, r_vec);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// L2: fused forward noising (diffusion.py:339-355) and reverse step (diffusion.py:291-326)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile) q_sample_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t,
                                                         const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1m_ac, int64_t T,
                                                         const float* __restrict__ cdf, const float* __restrict__ loc, uint64_t seed,
                                                         uint64_t rng_offset, uint64_t row_offset, float* __restrict__ x_t,
                                                         float* __restrict__ target3, float* __restrict__ noise, float* __restrict__ score3,
                                                         int64_t n, unsigned vecmask) {
  extern __shared__ float4 smem4[];
  float* s_in = reinterpret_cast<float*>(smem4);  // kTile*9  x0
  float* s_xt = s_in + kTile * 9;                 // kTile*9
  float* s_nz = s_xt + kTile * 9;                 // kTile*9  (noise, optional)
  float* s_tg = s_nz + kTile * 9;                 // kTile*3  (target)
  float* s_sc = s_tg + kTile * 3;                 // kTile*3  (score)
  float* s_loc = s_sc + kTile * 3;                // 1000
  for (int k = threadIdx.x; k < kCdf; k += kTile) s_loc[k] = loc[k];
  const int64_t tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int rows = (int)((n - row0) < kTile ? (n - row0) : kTile);
    tile_load<9>(s_in, x0 + row0 * 9, rows, vecmask & 1u);
    __syncthreads();
    const int r = threadIdx.x;
    if (r < rows) {
      const int64_t i = row0 + r;
      int64_t ti = t[i];
      ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
      const float eps = __ldg(sqrt_1m_ac + ti);
      const float sc = __ldg(sqrt_ac + ti);
      const NoiseDraw d = draw_axis_u(seed, row_offset + (uint64_t)i, rng_offset);
      const float ang = igso3_angle_from_uniform(cdf + ti * kCdf, s_loc, d.u);
      const Mat3 nz = rodrigues(d.axis, ang);
      const Mat3 x = sm_mat(s_in, r);
      sm_put_mat(s_xt, r, mul_nn(scale_rot(x, sc), nz));                      // diffusion.py:344-346
      if (noise) sm_put_mat(s_nz, r, nz);
      if (target3) {
        const float k = ang / eps;                                             // diffusion.py:355
        sm_put_vec(s_tg, r, Vec3{k * d.axis.x, k * d.axis.y, k * d.axis.z});
      }
      if (score3) {
        float lf, g;
        igso3_logf_g(ang, eps, kAuto, 2000, &lf, &g);
        sm_put_vec(s_sc, r, Vec3{g * d.axis.x, g * d.axis.y, g * d.axis.z});
      }
    }
    __syncthreads();
    tile_store<9>(x_t + row0 * 9, s_xt, rows, (vecmask >> 1) & 1u);
    if (noise) tile_store<9>(noise + row0 * 9, s_nz, rows, (vecmask >> 2) & 1u);
    if (target3) tile_store<3>(target3 + row0 * 3, s_tg, rows, (vecmask >> 3) & 1u);
    if (score3) tile_store<3>(score3 + row0 * 3, s_sc, rows, (vecmask >> 4) & 1u);
  }
}

struct QSampleGivenOp {  // diffusion.py:339-346 with noise supplied
  SO3D_OP_ARRAYS(2, 0, 1, 0)
  const int64_t* t;
  const float* sqrt_ac;
  int64_t T;
  __device__ void row(int64_t i, const Mat3* a9, const Vec3*, Mat3* o9, Vec3*) const {
    int64_t ti = t[i];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    o9[0] = mul_nn(scale_rot(a9[0], __ldg(sqrt_ac + ti)), a9[1]);
  }
};

// ------------------------------------------------------------------------------------------------
// Pipelined tile movement for the fused HBM-bound kernels (so3d_tma.cuh): persistent CTAs, two
// shared-memory stages per array, one elected thread driving the TMA engine.
//   iteration k (tile = blockIdx.x + k gridDim.x, stage = k & 1):
//     wait full[stage]  ->  every thread copies its row to registers  ->  barrier A (stage may be refilled;
//     thread 0 has drained the bulk store issued two iterations ago)  ->  thread 0 issues the bulk loads of
//     tile k+2 into `stage`  ->  arithmetic  ->  results to out[stage]  ->  proxy fence + barrier B  ->
//     thread 0 issues the bulk store of out[stage].
// A tile that is not a full, 16-byte aligned tile (ragged tail, unaligned views) is moved cooperatively
// with ordinary loads/stores at the same points of the schedule.
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void coop_load(float* __restrict__ sm, const float* __restrict__ g, int rows) {
  for (int i = threadIdx.x; i < rows * W; i += kTile) sm[i] = __ldcs(g + i);
}
template <int W>
__device__ __forceinline__ void coop_store(float* __restrict__ g, const float* __restrict__ sm, int rows) {
  for (int i = threadIdx.x; i < rows * W; i += kTile) __stcs(g + i, sm[i]);
}

// CDF row, grid angles and the guide of one shared table row, staged in shared memory by the whole CTA.
struct SharedCdf {
  float* loc;       // kCdf (+1 pad)
  float* trap;      // kCdf (+1 pad)
  uint16_t* guide;  // kGuideStride
};
__device__ __forceinline__ void stage_shared_cdf(const SharedCdf& sc, const float* __restrict__ cdf_row, const float* __restrict__ loc) {
  for (int k = threadIdx.x; k < kCdf; k += kTile) {
    sc.loc[k] = loc[k];
    sc.trap[k] = cdf_row[k];
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= kGuide; k += kTile)
    sc.guide[k] = (uint16_t)cdf_count_le(sc.trap, (float)k * (1.0f / (float)kGuide), 0, kCdf);
  __syncthreads();
}

// distributions.py:15-30 companion: guide[row][k] = #{j : trap[row][j] <= k/1024} (so3d_math.cuh).
__global__ void __launch_bounds__(kTile) cdf_guide_kernel(const float* __restrict__ cdf, uint16_t* __restrict__ guide) {
  __shared__ float s_trap[kGrid];
  const int64_t row = blockIdx.x;
  for (int k = threadIdx.x; k < kCdf; k += kTile) s_trap[k] = cdf[row * kCdf + k];
  __syncthreads();
  for (int k = threadIdx.x; k < kGuideStride; k += kTile)
    guide[row * kGuideStride + k] = (k <= kGuide) ? (uint16_t)cdf_count_le(s_trap, (float)k * (1.0f / (float)kGuide), 0, kCdf) : (uint16_t)kCdf;
}

// diffusion.py:291-326, fused reverse step.
//   kSharedT: the whole batch shares t (the reference's semantics, Q7): schedule scalars are uniform and the
//             posterior CDF row + guide live in shared memory; otherwise per-row t with table rows read from L2.
//   kX0:      also write x0_hat.
template <bool kSharedT, bool kX0>
__global__ void __launch_bounds__(kTile) p_step_kernel(const float* __restrict__ x_t, const float* __restrict__ pred3,
                                                       const int64_t* __restrict__ t, const float* __restrict__ recip,
                                                       const float* __restrict__ recipm1, const float* __restrict__ coef1,
                                                       const float* __restrict__ coef2, int64_t T, const float* __restrict__ post_cdf,
                                                       const uint16_t* __restrict__ post_guide, const float* __restrict__ loc,
                                                       uint64_t seed, uint64_t rng_offset, uint64_t row_offset, float* __restrict__ out,
                                                       float* __restrict__ x0_hat_out, int64_t n, int use_tma) {
  extern __shared__ float4 smem4[];
  float* s_x = reinterpret_cast<float*>(smem4);     // [2][kTile*9]
  float* s_p = s_x + 2 * kTile * 9;                 // [2][kTile*3]
  float* s_o = s_p + 2 * kTile * 3;                 // [2][kTile*9]
  float* s_h = s_o + 2 * kTile * 9;                 // [2][kTile*9] (kX0)
  float* s_tab = s_h + (kX0 ? 2 * kTile * 9 : 0);
  SharedCdf sc;
  sc.loc = s_tab;
  sc.trap = s_tab + kGrid;
  sc.guide = reinterpret_cast<uint16_t*>(s_tab + 2 * kGrid);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab + 2 * kGrid + kGuideStride / 2 + 1);  // 8-byte aligned: all counts even
  const int tid = threadIdx.x;

  int64_t t_shared = 0;
  if (kSharedT) {
    t_shared = t[0];
    t_shared = t_shared < 0 ? 0 : (t_shared >= T ? T - 1 : t_shared);
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (post_cdf) {
    if (kSharedT) {
      stage_shared_cdf(sc, post_cdf + t_shared * kCdf, loc);
    } else {
      for (int k = tid; k < kCdf; k += kTile) sc.loc[k] = loc[k];
    }
  }
  __syncthreads();

  const int64_t tiles = (n + kTile - 1) / kTile;
  const int64_t my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_rows = [&](int64_t k) -> int {
    const int64_t row0 = (blockIdx.x + k * gridDim.x) * kTile;
    return (int)((n - row0) < kTile ? (n - row0) : kTile);
  };
  auto issue_load = [&](int64_t k) {  // thread 0 only
    const int64_t row0 = (blockIdx.x + k * gridDim.x) * kTile;
    const int st = (int)(k & 1);
    mbar_expect_tx(&bars[st], kTile * 12 * sizeof(float));
    bulk_load(s_x + st * kTile * 9, x_t + row0 * 9, kTile * 9 * sizeof(float), &bars[st]);
    bulk_load(s_p + st * kTile * 3, pred3 + row0 * 3, kTile * 3 * sizeof(float), &bars[st]);
  };
  if (tid == 0 && use_tma) {
    if (my_tiles > 0 && tile_rows(0) == kTile) issue_load(0);
    if (my_tiles > 1 && tile_rows(1) == kTile) issue_load(1);
  }

  float k_recip = 0.f, k_recipm1 = 0.f, k_c1 = 0.f, k_c2 = 0.f;
  if (kSharedT) {
    k_recip = __ldg(recip + t_shared); k_recipm1 = __ldg(recipm1 + t_shared);
    k_c1 = __ldg(coef1 + t_shared); k_c2 = __ldg(coef2 + t_shared);
  }

  for (int64_t k = 0; k < my_tiles; ++k) {
    const int st = (int)(k & 1);
    const int64_t row0 = (blockIdx.x + k * gridDim.x) * kTile;
    const int rows = tile_rows(k);
    const bool tma = use_tma && rows == kTile;
    float* sx = s_x + st * kTile * 9;
    float* sp = s_p + st * kTile * 3;
    float* so = s_o + st * kTile * 9;
    float* sh = s_h + st * kTile * 9;
    if (tma) {
      mbar_wait(&bars[st], (uint32_t)((k >> 1) & 1));
    } else {
      coop_load<9>(sx, x_t + row0 * 9, rows);
      coop_load<3>(sp, pred3 + row0 * 3, rows);
      __syncthreads();
    }
    const Mat3 x = sm_mat(sx, tid);
    const Vec3 p = sm_vec(sp, tid);
    if (tid == 0) bulk_wait_read<1>();  // the store that read out[st] two iterations ago has drained
    __syncthreads();                    // A
    if (tid == 0 && use_tma && k + 2 < my_tiles && tile_rows(k + 2) == kTile) issue_load(k + 2);

    if (tid < rows) {
      const int64_t i = row0 + tid;
      int64_t ti = t_shared;
      if (!kSharedT) {
        ti = t[i];
        ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
        k_recip = __ldg(recip + ti); k_recipm1 = __ldg(recipm1 + ti);
        k_c1 = __ldg(coef1 + ti); k_c2 = __ldg(coef2 + ti);
      }
      Quat qh;
      Quat qm = p_mean_quat(x, p, k_recip, k_recipm1, k_c1, k_c2, &qh);
      if (post_cdf && ti != 0) {                                                   // diffusion.py:320-326
        const NoiseDraw d = draw_axis_u(seed, row_offset + (uint64_t)i, rng_offset);
        float ang;
        if (kSharedT) ang = igso3_angle_from_uniform_guided(sc.trap, sc.loc, sc.guide, d.u);
        else if (post_guide) ang = igso3_angle_from_uniform_guided(post_cdf + ti * kCdf, sc.loc, post_guide + ti * kGuideStride, d.u);
        else ang = igso3_angle_from_uniform(post_cdf + ti * kCdf, sc.loc, d.u);
        qm = qmul(qm, quat_axis_angle(d.axis, ang));
      }
      sm_put_mat(so, tid, quat_to_mat_unit(qm));
      if (kX0) sm_put_mat(sh, tid, quat_to_mat_unit(qh));
    }
    if (tma) {
      fence_proxy_async();
      __syncthreads();  // B
      if (tid == 0) {
        bulk_store(out + row0 * 9, so, kTile * 9 * sizeof(float));
        if (kX0) bulk_store(x0_hat_out + row0 * 9, sh, kTile * 9 * sizeof(float));
        bulk_commit();
      }
    } else {
      __syncthreads();  // B
      coop_store<9>(out + row0 * 9, so, rows);
      if (kX0) coop_store<9>(x0_hat_out + row0 * 9, sh, rows);
    }
  }
  if (tid == 0) bulk_wait_read<0>();
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int so3d_version(void) { return SO3D_VERSION; }
const char* so3d_last_error(void) { return g_err; }

#define SO3D_REQUIRE(cond, msg) \
  if (!(cond)) return fail(SO3D_EINVAL, msg)

int so3d_log_f32(const float* R, float* out9, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && out9), "so3d_log_f32: null pointer");
  LogOp op;
  op.in9[0] = R; op.out9[0] = out9;
  return launch_rowwise(op, n, stream, "so3d_log_f32");
}

int so3d_logvec_f32(const float* R, float* out3, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && out3), "so3d_logvec_f32: null pointer");
  LogVecOp op;
  op.in9[0] = R; op.out3[0] = out3;
  return launch_rowwise(op, n, stream, "so3d_logvec_f32");
}

int so3d_rmat_to_aa_f32(const float* R, float* axis3, float* angle, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && axis3 && angle), "so3d_rmat_to_aa_f32: null pointer");
  RmatToAaOp op;
  op.in9[0] = R; op.out3[0] = axis3; op.angle = angle;
  return launch_rowwise(op, n, stream, "so3d_rmat_to_aa_f32");
}

int so3d_aa_to_rmat_f32(const float* axis3, const float* angle, float* R, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (axis3 && angle && R), "so3d_aa_to_rmat_f32: null pointer");
  AaToRmatOp op;
  op.in3[0] = axis3; op.angle = angle; op.out9[0] = R;
  return launch_rowwise(op, n, stream, "so3d_aa_to_rmat_f32");
}

int so3d_expvec_f32(const float* v3, float* R, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (v3 && R), "so3d_expvec_f32: null pointer");
  ExpVecOp op;
  op.in3[0] = v3; op.out9[0] = R;
  return launch_rowwise(op, n, stream, "so3d_expvec_f32");
}

int so3d_scale_f32(const float* R, const float* s, int s_stride, float* out, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && s && out), "so3d_scale_f32: null pointer");
  SO3D_REQUIRE(s_stride == 0 || s_stride == 1, "so3d_scale_f32: s_stride must be 0 or 1");
  ScaleOp op;
  op.in9[0] = R; op.s = s; op.s_stride = s_stride; op.out9[0] = out;
  return launch_rowwise(op, n, stream, "so3d_scale_f32");
}

int so3d_quat_to_rmat_f32(const float* q4, float* R, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (q4 && R), "so3d_quat_to_rmat_f32: null pointer");
  QuatToRmatOp op;
  op.q = q4; op.q_vec = aligned16(q4); op.out9[0] = R;
  return launch_rowwise(op, n, stream, "so3d_quat_to_rmat_f32");
}

int so3d_rmat_to_quat_f32(const float* R, float* q4, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && q4), "so3d_rmat_to_quat_f32: null pointer");
  SO3D_REQUIRE(aligned16(q4), "so3d_rmat_to_quat_f32: q4 must be 16-byte aligned");
  RmatToQuatOp op;
  op.in9[0] = R; op.q = q4;
  return launch_rowwise(op, n, stream, "so3d_rmat_to_quat_f32");
}

}  // extern "C"

template <bool TA, bool TB>
static int compose_dispatch(const float* A, int a_stride, const float* B, int b_stride, float* C, int64_t n, void* stream) {
  if (a_stride && b_stride) {
    ComposeOp<TA, TB> op;
    op.in9[0] = A; op.in9[1] = B; op.out9[0] = C;
    return launch_rowwise(op, n, stream, "so3d_compose_f32");
  } else if (!a_stride && b_stride) {
    ComposeSharedOp<TA, TB, true> op;
    op.in9[0] = B; op.shared = A; op.out9[0] = C;
    return launch_rowwise(op, n, stream, "so3d_compose_f32");
  } else if (a_stride && !b_stride) {
    ComposeSharedOp<TA, TB, false> op;
    op.in9[0] = A; op.shared = B; op.out9[0] = C;
    return launch_rowwise(op, n, stream, "so3d_compose_f32");
  }
  return fail(SO3D_EINVAL, "so3d_compose_f32: at most one operand may be shared");
}

extern "C" {

int so3d_compose_f32(const float* A, int a_stride, int trans_a, const float* B, int b_stride, int trans_b, float* C,
                     int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (A && B && C), "so3d_compose_f32: null pointer");
  if (trans_a && trans_b) return compose_dispatch<true, true>(A, a_stride, B, b_stride, C, n, stream);
  if (trans_a) return compose_dispatch<true, false>(A, a_stride, B, b_stride, C, n, stream);
  if (trans_b) return compose_dispatch<false, true>(A, a_stride, B, b_stride, C, n, stream);
  return compose_dispatch<false, false>(A, a_stride, B, b_stride, C, n, stream);
}

int so3d_rmat_dist_f32(const float* A, const float* B, float* out, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (A && B && out), "so3d_rmat_dist_f32: null pointer");
  RmatDistOp op;
  op.in9[0] = A; op.in9[1] = B; op.out = out;
  return launch_rowwise(op, n, stream, "so3d_rmat_dist_f32");
}

int so3d_lerp_f32(const float* A, const float* B, const float* w, int w_stride, float* out, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (A && B && w && out), "so3d_lerp_f32: null pointer");
  SO3D_REQUIRE(w_stride == 0 || w_stride == 1, "so3d_lerp_f32: w_stride must be 0 or 1");
  LerpOp op;
  op.in9[0] = A; op.in9[1] = B; op.w = w; op.w_stride = w_stride; op.out9[0] = out;
  return launch_rowwise(op, n, stream, "so3d_lerp_f32");
}

int so3d_log_bwd_f32(const float* R, const float* G9, float* gR, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && G9 && gR), "so3d_log_bwd_f32: null pointer");
  LogBwdOp op;
  op.in9[0] = R; op.in9[1] = G9; op.out9[0] = gR;
  return launch_rowwise(op, n, stream, "so3d_log_bwd_f32");
}

int so3d_aa_to_rmat_bwd_f32(const float* axis3, const float* angle, const float* G9, float* g_axis3, float* g_angle,
                            int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (axis3 && angle && G9 && g_axis3 && g_angle), "so3d_aa_to_rmat_bwd_f32: null pointer");
  AaToRmatBwdOp op;
  op.in9[0] = G9; op.in3[0] = axis3; op.angle = angle; op.out3[0] = g_axis3; op.g_angle = g_angle;
  return launch_rowwise(op, n, stream, "so3d_aa_to_rmat_bwd_f32");
}

int so3d_expvec_bwd_f32(const float* v3, const float* G9, float* g_v3, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (v3 && G9 && g_v3), "so3d_expvec_bwd_f32: null pointer");
  ExpVecBwdOp op;
  op.in9[0] = G9; op.in3[0] = v3; op.out3[0] = g_v3;
  return launch_rowwise(op, n, stream, "so3d_expvec_bwd_f32");
}

int so3d_scale_bwd_f32(const float* R, const float* s, int s_stride, const float* G9, float* gR, float* g_s, int64_t n,
                       void* stream) {
  SO3D_REQUIRE(n == 0 || (R && s && G9 && gR), "so3d_scale_bwd_f32: null pointer");
  SO3D_REQUIRE(s_stride == 0 || s_stride == 1, "so3d_scale_bwd_f32: s_stride must be 0 or 1");
  ScaleBwdOp op;
  op.in9[0] = R; op.in9[1] = G9; op.s = s; op.s_stride = s_stride; op.g_s = g_s; op.out9[0] = gR;
  return launch_rowwise(op, n, stream, "so3d_scale_bwd_f32");
}

static int check_mode(int mode, int L) {
  if (mode < 0 || mode > 3) return fail(SO3D_EINVAL, "mode must be SO3D_MODE_{SERIES,CLOSED,AUTO,SERIES_ADAPTIVE}");
  if (mode != SO3D_MODE_CLOSED && (L < 1 || L > 2896)) return fail(SO3D_EINVAL, "series truncation L must be in [1, 2896]");
  return 0;
}

int so3d_igso3_density_f32(const float* omega, const float* eps, int eps_stride, float* f, int64_t n, int mode, int L,
                           void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(omega && eps && f, "so3d_igso3_density_f32: null pointer");
  SO3D_REQUIRE(eps_stride == 0 || eps_stride == 1, "eps_stride must be 0 or 1");
  if (int rc = check_mode(mode, L)) return rc;
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  density_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(omega, eps, eps_stride, f, n, mode, L);
  return check_launch("so3d_igso3_density_f32");
}

int so3d_igso3_logp_score_f32(const float* R, const float* eps, int eps_stride, float* logp, float* score3, float* dlogf,
                              int64_t n, int mode, int L, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && eps && logp), "so3d_igso3_logp_score_f32: null pointer");
  SO3D_REQUIRE(eps_stride == 0 || eps_stride == 1, "eps_stride must be 0 or 1");
  if (int rc = check_mode(mode, L)) return rc;
  LogpScoreOp op;
  op.in9[0] = R; op.eps = eps; op.eps_stride = eps_stride; op.logp = logp; op.dlogf = dlogf; op.out3[0] = score3;
  op.mode = mode; op.L = L;
  return launch_rowwise(op, n, stream, "so3d_igso3_logp_score_f32");
}

int so3d_igso3_logp_bwd_f32(const float* R, const float* dlogf, const float* gout, float* gR, int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (R && dlogf && gout && gR), "so3d_igso3_logp_bwd_f32: null pointer");
  LogpBwdOp op;
  op.in9[0] = R; op.dlogf = dlogf; op.gout = gout; op.out9[0] = gR;
  return launch_rowwise(op, n, stream, "so3d_igso3_logp_bwd_f32");
}

int so3d_igso3_cdf_table_f32(const float* eps, int64_t rows, const float* grid_loc, const float* haar_w, float* trap_out,
                             int quirks, void* stream) {
  SO3D_REQUIRE(rows >= 0, "negative rows");
  if (rows == 0) return 0;
  SO3D_REQUIRE(eps && grid_loc && haar_w && trap_out, "so3d_igso3_cdf_table_f32: null pointer");
  SO3D_REQUIRE(rows <= 0x7fffffff, "so3d_igso3_cdf_table_f32: too many rows");
  cdf_table_kernel<<<(int)rows, 256, 0, (cudaStream_t)stream>>>(eps, grid_loc, haar_w, trap_out, quirks);
  return check_launch("so3d_igso3_cdf_table_f32");
}

int so3d_igso3_sample_f32(const float* cdf, const float* loc, int64_t rows, const int64_t* row_idx, int64_t row,
                          const float* u, const float* axes3, uint64_t seed, uint64_t rng_offset, uint64_t row_offset,
                          const float* mean, int mean_stride, float* R, float* angle, float* axis3, int64_t n, void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(cdf && loc && R, "so3d_igso3_sample_f32: null pointer");
  SO3D_REQUIRE(rows > 0, "so3d_igso3_sample_f32: empty table");
  SO3D_REQUIRE(row_idx || (row >= 0 && row < rows), "so3d_igso3_sample_f32: row out of range");
  SO3D_REQUIRE(mean_stride == 0 || mean_stride == 1, "mean_stride must be 0 or 1");
  const size_t smem = sizeof(float) * (kTile * 9 + 2000);
  const int grid = grid_for(n, 8);
  if (row_idx)
    sample_kernel<false><<<grid, kTile, smem, (cudaStream_t)stream>>>(cdf, loc, row_idx, 0, rows, u, axes3, seed, rng_offset, row_offset,
                                                                       mean, mean_stride, R, angle, axis3, n, aligned16(R));
  else
    sample_kernel<true><<<grid, kTile, smem, (cudaStream_t)stream>>>(cdf, loc, nullptr, row, rows, u, axes3, seed, rng_offset, row_offset,
                                                                      mean, mean_stride, R, angle, axis3, n, aligned16(R));
  return check_launch("so3d_igso3_sample_f32");
}

int so3d_q_sample_f32(const float* x0, const int64_t* t, const float* sqrt_ac, const float* sqrt_1m_ac, int64_t T,
                      const float* cdf, const float* loc, uint64_t seed, uint64_t rng_offset, uint64_t row_offset, float* x_t,
                      float* target3, float* noise, float* score3, int64_t n, void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(x0 && t && sqrt_ac && sqrt_1m_ac && cdf && loc && x_t, "so3d_q_sample_f32: null pointer");
  SO3D_REQUIRE(T > 0, "so3d_q_sample_f32: T must be positive");
  const unsigned mask = (aligned16(x0) ? 1u : 0u) | (aligned16(x_t) ? 2u : 0u) | (aligned16(noise) ? 4u : 0u) |
                        (aligned16(target3) ? 8u : 0u) | (aligned16(score3) ? 16u : 0u);
  const size_t smem = sizeof(float) * (kTile * (9 * 3 + 3 * 2) + 1000);
  q_sample_kernel<<<grid_for(n, 6), kTile, smem, (cudaStream_t)stream>>>(x0, t, sqrt_ac, sqrt_1m_ac, T, cdf, loc, seed, rng_offset,
                                                                          row_offset, x_t, target3, noise, score3, n, mask);
  return check_launch("so3d_q_sample_f32");
}

int so3d_q_sample_given_f32(const float* x0, const int64_t* t, const float* sqrt_ac, int64_t T, const float* noise, float* x_t,
                            int64_t n, void* stream) {
  SO3D_REQUIRE(n == 0 || (x0 && t && sqrt_ac && noise && x_t), "so3d_q_sample_given_f32: null pointer");
  SO3D_REQUIRE(T > 0, "so3d_q_sample_given_f32: T must be positive");
  QSampleGivenOp op;
  op.in9[0] = x0; op.in9[1] = noise; op.out9[0] = x_t; op.t = t; op.sqrt_ac = sqrt_ac; op.T = T;
  return launch_rowwise(op, n, stream, "so3d_q_sample_given_f32");
}

int so3d_igso3_cdf_guide_u16(const float* cdf, int64_t rows, uint16_t* guide_out, void* stream) {
  SO3D_REQUIRE(rows >= 0, "negative rows");
  if (rows == 0) return 0;
  SO3D_REQUIRE(cdf && guide_out, "so3d_igso3_cdf_guide_u16: null pointer");
  SO3D_REQUIRE(rows <= 0x7fffffff, "so3d_igso3_cdf_guide_u16: too many rows");
  cdf_guide_kernel<<<(int)rows, kTile, 0, (cudaStream_t)stream>>>(cdf, guide_out);
  return check_launch("so3d_igso3_cdf_guide_u16");
}

}  // extern "C"

template <bool kSharedT, bool kX0>
static int launch_p_step(const float* x_t, const float* pred3, const int64_t* t, const float* recip, const float* recipm1,
                         const float* coef1, const float* coef2, int64_t T, const float* post_cdf, const uint16_t* post_guide,
                         const float* loc, uint64_t seed, uint64_t rng_offset, uint64_t row_offset, float* out, float* x0_hat_out,
                         int64_t n, void* stream) {
  const size_t smem = sizeof(float) * (size_t)(2 * kTile * (9 + 3 + 9 + (kX0 ? 9 : 0)) + 2 * kGrid + kGuideStride / 2 + 1) + 2 * sizeof(uint64_t);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(p_step_kernel<kSharedT, kX0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  const int use_tma = aligned16(x_t) && aligned16(pred3) && aligned16(out) && (!kX0 || aligned16(x0_hat_out));
  p_step_kernel<kSharedT, kX0><<<grid_for(n, kX0 ? 3 : 4), kTile, smem, (cudaStream_t)stream>>>(
      x_t, pred3, t, recip, recipm1, coef1, coef2, T, post_cdf, post_guide, loc, seed, rng_offset, row_offset, out, x0_hat_out, n, use_tma);
  return check_launch("so3d_p_sample_f32");
}

extern "C" {

int so3d_p_sample_f32(const float* x_t, const float* pred3, const int64_t* t, int t_stride, const float* recip,
                      const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                      const uint16_t* post_guide, const float* loc, uint64_t seed, uint64_t rng_offset, uint64_t row_offset,
                      float* out, float* x0_hat_out, int64_t n, void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(x_t && pred3 && t && recip && recipm1 && coef1 && coef2 && out, "so3d_p_sample_f32: null pointer");
  SO3D_REQUIRE(T > 0, "so3d_p_sample_f32: T must be positive");
  SO3D_REQUIRE(t_stride == 0 || t_stride == 1, "t_stride must be 0 or 1");
  SO3D_REQUIRE(!post_cdf || loc, "so3d_p_sample_f32: loc required with post_cdf");
#define SO3D_PSTEP(S, X) \
  launch_p_step<S, X>(x_t, pred3, t, recip, recipm1, coef1, coef2, T, post_cdf, post_guide, loc, seed, rng_offset, row_offset, out, x0_hat_out, n, stream)
  if (t_stride == 0) return x0_hat_out ? SO3D_PSTEP(true, true) : SO3D_PSTEP(true, false);
  return x0_hat_out ? SO3D_PSTEP(false, true) : SO3D_PSTEP(false, false);
#undef SO3D_PSTEP
}

}  // extern "C"
