// so3d_loop.cu -- the whole reverse process of diffusion.py:328-337 in ONE launch when nothing outside the manifold
// step happens between two steps (no denoiser: pred = 0, or a fixed per-particle prediction): BASELINE configs[2],
// "1000 steps x 2^24 particles".
//
// Particles are independent and the step's noise is a pure function of (seed, global row, step), so the loop over the
// steps can be turned inside out: a CTA takes a chunk of particles, keeps their 3x3 matrices RESIDENT IN SHARED MEMORY
// and runs all T steps on them -- x_T is read from HBM once, x_0 written once, and between the steps there is no
// launch, no grid-wide synchronisation and not even a CTA barrier (every thread only ever touches its own rows).  The
// per-step launches of so3d_p_sample_f32 move 84 B per particle-step through HBM and are instruction-issue bound at
// 0.60 of the HBM roofline; this kernel has no HBM traffic to speak of (72 B per particle per 1000 steps) and is bound
// by instruction issue alone.  What a step still reads from memory is the step's four schedule scalars and, per
// particle, ONE 16-byte guide record of the posterior CDF row (L2-resident table, see so3d_math.cuh), requested one
// particle ahead so that its latency is covered by the arithmetic of the current particle.
//
// Bit-identical to the sequence of launches
//     for t = t_hi .. t_lo:  so3d_p_sample_f32(x, pred, t (shared), ..., seed, rng_offset0 + t, row_offset, x')
// (the matrix is rounded to float32 between the steps exactly as the stores / loads of those launches do).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/so3d.h"
#include "so3d_common.cuh"
#include "so3d_math.cuh"
#include "so3d_lanes.cuh"

using namespace so3d;

namespace {

constexpr int kLT = 256;       // threads per CTA; a tile = 256 rows, thread `tid` owns row `tid` of every resident tile
constexpr int kLoopTiles = 5;  // resident tiles per CTA: 5 x 9216 B + the grid locations = 50 KB -> 4 CTAs (32 warps) per SM

struct LoopArgs {
  const float* x;
  const float* pred;
  float* out;
  int64_t n, rows_per_cta;
  int64_t t_hi, t_lo;
  const float* recip;
  const float* recipm1;
  const float* coef1;
  const float* coef2;
  const float* post_cdf;
  const uint32_t* post_guide;
  const float* loc;
  PhiloxRoundKeys keys;
  uint64_t rng_offset0, row_offset;
};

struct Pre {  // what a row needs from the RNG and the guide table for one step, fetched one row ahead
  NoiseDraw d;
  uint4 rec;
};

template <bool kHasPred>
__global__ void __launch_bounds__(kLT, 4) p_sample_loop_kernel(const LoopArgs a) {
  extern __shared__ float4 smem4[];
  float* s_rows = reinterpret_cast<float*>(smem4);
  float* s_loc = s_rows + kLoopTiles * kLT * 9;
  const int tid = threadIdx.x;
  for (int k = tid; k < kCdf; k += kLT) s_loc[k] = a.loc[k];

  const int64_t cta_lo = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t cta_hi = cta_lo + a.rows_per_cta < a.n ? cta_lo + a.rows_per_cta : a.n;
  for (int64_t pass_lo = cta_lo; pass_lo < cta_hi; pass_lo += kLoopTiles * kLT) {
    const int64_t left = cta_hi - pass_lo;
    const int rows = (int)(left < kLoopTiles * kLT ? left : kLoopTiles * kLT);
    __syncthreads();  // the previous pass's write-back has finished reading the rows (first pass: s_loc is staged)
    {
      const float* __restrict__ src = a.x + pass_lo * 9;
      for (int i = tid; i < rows * 9; i += kLT) s_rows[i] = __ldcs(src + i);
    }
    __syncthreads();
    const int my = tid < rows ? (rows - tid + kLT - 1) / kLT : 0;  // this thread's rows: tid, tid + 256, ...

    auto prefetch = [&](int64_t step, int k) -> Pre {
      Pre p;
      p.d = NoiseDraw{Vec3{0.f, 0.f, 1.f}, 0.f};
      p.rec = make_uint4(0, 0, 0, 0);
      if (step != 0) {  // diffusion.py:320: no noise at t == 0
        const uint64_t row = a.row_offset + (uint64_t)(pass_lo + (int64_t)k * kLT + tid);
        p.d = draw_axis_u(a.keys, row, a.rng_offset0 + (uint64_t)step);
        p.rec = __ldg(reinterpret_cast<const uint4*>(a.post_guide) + step * kGuideRecs + guide_bucket(p.d.u));
      }
      return p;
    };

    Pre cur;
    if (my > 0) cur = prefetch(a.t_hi, 0);
    for (int64_t step = a.t_hi; step >= a.t_lo && my > 0; --step) {
      const float k_recip = __ldg(a.recip + step), k_recipm1 = __ldg(a.recipm1 + step);
      const float k_c1 = __ldg(a.coef1 + step), k_c2 = __ldg(a.coef2 + step);
      const float* __restrict__ trap = a.post_cdf + step * kCdf;
#pragma unroll 1
      for (int k = 0; k < my; ++k) {
        int nk = k + 1;
        int64_t ns = step;
        if (nk == my) {
          nk = 0;
          ns = step - 1;
        }
        Pre nxt = cur;
        if (ns >= a.t_lo) nxt = prefetch(ns, nk);  // its guide record lands while this row is computed

        float* rp = s_rows + (k * kLT + tid) * 9;
        Mat3 x;
#pragma unroll
        for (int j = 0; j < 9; ++j) x.m[j] = rp[j];
        Quat qh;
        Quat qm;
        if (kHasPred) {
          const float* pp = a.pred + (pass_lo + (int64_t)k * kLT + tid) * 3;
          qm = p_mean_quat(x, Vec3{__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)}, k_recip, k_recipm1, k_c1, k_c2, &qh);
        } else {
          qm = p_mean_quat_nopred(x, k_recip, k_c1, k_c2, &qh);
        }
        if (step != 0) {
          const GuideRec rec{cur.rec.x, __uint_as_float(cur.rec.y), __uint_as_float(cur.rec.z), __uint_as_float(cur.rec.w)};
          const float ang = igso3_angle_from_record(trap, s_loc, rec, cur.d.u);
          qm = qmul(qm, quat_axis_angle(cur.d.axis, ang));
        }
        const Mat3 o = quat_to_mat_unit(qm);
#pragma unroll
        for (int j = 0; j < 9; ++j) rp[j] = o.m[j];
        cur = nxt;
      }
    }
    __syncthreads();
    {
      float* __restrict__ dst = a.out + pass_lo * 9;
      for (int i = tid; i < rows * 9; i += kLT) __stcs(dst + i, s_rows[i]);
    }
  }
}

// Two rows per thread, FP32 work packed (so3d_lanes.cuh): thread `tid` owns row `tid` of every resident tile and walks the
// tiles in PAIRS -- rows (k, tid) and (k + 1, tid) ride in the two lanes of FFMA2 / FMUL2 / FADD2 instructions, which halves the
// issue slots of the ~45 % of the step that is FP32 arithmetic (the step is issue-bound).  Same IEEE operations per lane as the
// one-row kernel, hence the same bits.  A trailing unpaired row takes the one-lane instantiation.
#ifndef SO3D_LOOP2_MINCTAS
#define SO3D_LOOP2_MINCTAS 2   // resident CTAs promised to ptxas (2: <= 128 registers, no spills; 3: <= 85)
#endif
#ifndef SO3D_LOOP2_TILES
#define SO3D_LOOP2_TILES 10    // resident tiles per CTA (even): 10 x 9216 B + loc = 96 KB -> 2 CTAs per SM; 6 -> 3 CTAs
#endif
constexpr int kLoopTiles2 = SO3D_LOOP2_TILES;

template <class L>
struct PreL;
template <>
struct PreL<L1> {
  Vec3L<L1> axis;
  float u[1];
  uint4 rec[1];
};
template <>
struct PreL<L2> {
  Vec3L<L2> axis;
  float u[2];
  uint4 rec[2];
};

template <bool kHasPred>
__global__ void __launch_bounds__(kLT, SO3D_LOOP2_MINCTAS) p_sample_loop_kernel2(const LoopArgs a) {
  extern __shared__ float4 smem4[];
  float* s_rows = reinterpret_cast<float*>(smem4);
  float* s_loc = s_rows + kLoopTiles2 * kLT * 9;
  const int tid = threadIdx.x;
  for (int k = tid; k < kCdf; k += kLT) s_loc[k] = a.loc[k];

  const int64_t cta_lo = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t cta_hi = cta_lo + a.rows_per_cta < a.n ? cta_lo + a.rows_per_cta : a.n;
  for (int64_t pass_lo = cta_lo; pass_lo < cta_hi; pass_lo += kLoopTiles2 * kLT) {
    const int64_t left = cta_hi - pass_lo;
    const int rows = (int)(left < kLoopTiles2 * kLT ? left : kLoopTiles2 * kLT);
    __syncthreads();
    {
      const float* __restrict__ src = a.x + pass_lo * 9;
      for (int i = tid; i < rows * 9; i += kLT) s_rows[i] = __ldcs(src + i);
    }
    __syncthreads();
    const int my = tid < rows ? (rows - tid + kLT - 1) / kLT : 0;  // this thread's rows: tid, tid + 256, ...
    const int pairs = my >> 1;

    auto draw2 = [&](int64_t step, int k) -> PreL<L2> {
      PreL<L2> p;
      const uint64_t row = a.row_offset + (uint64_t)(pass_lo + (int64_t)k * kLT + tid);
      const uint64_t off = a.rng_offset0 + (uint64_t)step;
      const U4 r0 = philox4x32_10(a.keys, row, off), r1 = philox4x32_10(a.keys, row + kLT, off);
      p.axis = sphere_from_uniforms_l(L2{u01(r0.x), u01(r1.x)}, L2{u01(r0.y), u01(r1.y)});
      p.u[0] = u01(r0.z);
      p.u[1] = u01(r1.z);
      const uint4* g = reinterpret_cast<const uint4*>(a.post_guide) + step * kGuideRecs;
      p.rec[0] = __ldg(g + guide_bucket(p.u[0]));
      p.rec[1] = __ldg(g + guide_bucket(p.u[1]));
      return p;
    };

    for (int64_t step = a.t_hi; step >= a.t_lo && my > 0; --step) {
      const float k_recip = __ldg(a.recip + step), k_recipm1 = __ldg(a.recipm1 + step);
      const float k_c1 = __ldg(a.coef1 + step), k_c2 = __ldg(a.coef2 + step);
      const float* __restrict__ trap = a.post_cdf + step * kCdf;
      PreL<L2> cur;
      if (pairs > 0 && step != 0) cur = draw2(step, 0);
#pragma unroll 1
      for (int pr = 0; pr < pairs; ++pr) {
        const int k = 2 * pr;
        PreL<L2> nxt = cur;
        if (pr + 1 < pairs && step != 0) nxt = draw2(step, k + 2);  // its guide records land while this pair is computed
        float* rp0 = s_rows + (k * kLT + tid) * 9;
        float* rp1 = rp0 + kLT * 9;
        Mat3L<L2> x;
#pragma unroll
        for (int j = 0; j < 9; ++j) x.m[j] = L2{rp0[j], rp1[j]};
        Vec3L<L2> pred{L2{0.f, 0.f}, L2{0.f, 0.f}, L2{0.f, 0.f}};
        if (kHasPred) {
          const float* pp = a.pred + (pass_lo + (int64_t)k * kLT + tid) * 3;
          const float* pq = pp + kLT * 3;
          pred = Vec3L<L2>{L2{__ldg(pp), __ldg(pq)}, L2{__ldg(pp + 1), __ldg(pq + 1)}, L2{__ldg(pp + 2), __ldg(pq + 2)}};
        }
        QuatL<L2> qh;
        QuatL<L2> qm = p_mean_quat_l<L2, kHasPred>(x, pred, k_recip, k_recipm1, k_c1, k_c2, &qh);
        if (step != 0) {
          const GuideRec g0{cur.rec[0].x, __uint_as_float(cur.rec[0].y), __uint_as_float(cur.rec[0].z), __uint_as_float(cur.rec[0].w)};
          const GuideRec g1{cur.rec[1].x, __uint_as_float(cur.rec[1].y), __uint_as_float(cur.rec[1].z), __uint_as_float(cur.rec[1].w)};
          const L2 ang{igso3_angle_from_record(trap, s_loc, g0, cur.u[0]), igso3_angle_from_record(trap, s_loc, g1, cur.u[1])};
          qm = qmul_l(qm, quat_axis_angle_l(cur.axis, ang));
        }
        const Mat3L<L2> o = quat_to_mat_unit_l(qm);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          rp0[j] = o.m[j].x;
          rp1[j] = o.m[j].y;
        }
        cur = nxt;
      }
      if (my & 1) {  // the unpaired last row: one lane, same operations
        const int k = my - 1;
        float* rp = s_rows + (k * kLT + tid) * 9;
        Mat3L<L1> x;
#pragma unroll
        for (int j = 0; j < 9; ++j) x.m[j] = L1{rp[j]};
        Vec3L<L1> pred{L1{0.f}, L1{0.f}, L1{0.f}};
        if (kHasPred) {
          const float* pp = a.pred + (pass_lo + (int64_t)k * kLT + tid) * 3;
          pred = Vec3L<L1>{L1{__ldg(pp)}, L1{__ldg(pp + 1)}, L1{__ldg(pp + 2)}};
        }
        QuatL<L1> qh;
        QuatL<L1> qm = p_mean_quat_l<L1, kHasPred>(x, pred, k_recip, k_recipm1, k_c1, k_c2, &qh);
        if (step != 0) {
          const uint64_t row = a.row_offset + (uint64_t)(pass_lo + (int64_t)k * kLT + tid);
          const U4 r0 = philox4x32_10(a.keys, row, a.rng_offset0 + (uint64_t)step);
          const Vec3L<L1> axis = sphere_from_uniforms_l(L1{u01(r0.x)}, L1{u01(r0.y)});
          const float u = u01(r0.z);
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(a.post_guide) + step * kGuideRecs + guide_bucket(u));
          const GuideRec g0{w.x, __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w)};
          qm = qmul_l(qm, quat_axis_angle_l(axis, L1{igso3_angle_from_record(trap, s_loc, g0, u)}));
        }
        const Mat3L<L1> o = quat_to_mat_unit_l(qm);
#pragma unroll
        for (int j = 0; j < 9; ++j) rp[j] = o.m[j].x;
      }
    }
    __syncthreads();
    {
      float* __restrict__ dst = a.out + pass_lo * 9;
      for (int i = tid; i < rows * 9; i += kLT) __stcs(dst + i, s_rows[i]);
    }
  }
}

template <bool kHasPred>
int launch_loop2(const LoopArgs& a, int grid, void* stream) {
  constexpr size_t smem = sizeof(float) * ((size_t)kLoopTiles2 * kLT * 9 + kGrid);
  static bool configured[64] = {};
  const int dev = so3d_host::current_device();
  if (!configured[dev]) {
    cudaFuncSetAttribute(p_sample_loop_kernel2<kHasPred>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  p_sample_loop_kernel2<kHasPred><<<grid, kLT, smem, (cudaStream_t)stream>>>(a);
  return so3d_host::check_launch("so3d_p_sample_loop_f32");
}

template <bool kHasPred>
int launch_loop(const LoopArgs& a, int grid, void* stream) {
  constexpr size_t smem = sizeof(float) * ((size_t)kLoopTiles * kLT * 9 + kGrid);
  static bool configured[64] = {};
  const int dev = so3d_host::current_device();
  if (!configured[dev]) {
    cudaFuncSetAttribute(p_sample_loop_kernel<kHasPred>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  p_sample_loop_kernel<kHasPred><<<grid, kLT, smem, (cudaStream_t)stream>>>(a);
  return so3d_host::check_launch("so3d_p_sample_loop_f32");
}

}  // namespace

extern "C" int so3d_p_sample_loop_f32(const float* x_t, const float* pred3, int64_t t_hi, int64_t t_lo, const float* recip,
                                      const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                                      const uint32_t* post_guide, const float* loc, uint64_t seed, uint64_t rng_offset0,
                                      uint64_t row_offset, float* out, int64_t n, void* stream) {
  if (n < 0) return so3d_host::fail(SO3D_EINVAL, "negative n");
  if (n == 0) return 0;
  if (!(x_t && recip && recipm1 && coef1 && coef2 && post_cdf && post_guide && loc && out))
    return so3d_host::fail(SO3D_EINVAL, "so3d_p_sample_loop_f32: null pointer");
  if (T <= 0 || t_lo < 0 || t_hi >= T || t_lo > t_hi) return so3d_host::fail(SO3D_EINVAL, "so3d_p_sample_loop_f32: need 0 <= t_lo <= t_hi < T");
  if ((reinterpret_cast<uintptr_t>(post_guide) & 15u) != 0) return so3d_host::fail(SO3D_EINVAL, "so3d_p_sample_loop_f32: post_guide must be 16-byte aligned");
  LoopArgs a;
  a.x = x_t; a.pred = pred3; a.out = out; a.n = n; a.t_hi = t_hi; a.t_lo = t_lo;
  a.recip = recip; a.recipm1 = recipm1; a.coef1 = coef1; a.coef2 = coef2;
  a.post_cdf = post_cdf; a.post_guide = post_guide; a.loc = loc;
  a.keys = make_philox_round_keys(seed); a.rng_offset0 = rng_offset0; a.row_offset = row_offset;
  // every CTA gets the same number of rows (a multiple of the tile): CTA time is proportional to rows, not to passes
  static const int one_lane = [] {  // A/B aid: SO3D_LOOP_LANES=1 selects the one-row-per-thread kernel
    const char* e = getenv("SO3D_LOOP_LANES");
    return (e && atoi(e) == 1) ? 1 : 0;
  }();
  const int ctas_per_sm = one_lane ? 4 : SO3D_LOOP2_MINCTAS;
  const int64_t tiles = (n + kLT - 1) / kLT;
  const int64_t cap = (int64_t)so3d_host::sm_count() * ctas_per_sm;
  const int64_t grid0 = tiles < cap ? tiles : cap;
  int64_t tiles_per_cta = (tiles + grid0 - 1) / grid0;
  if (!one_lane && tiles_per_cta > 1) tiles_per_cta += tiles_per_cta & 1;  // even: every thread's rows pair up
  a.rows_per_cta = tiles_per_cta * kLT;
  const int grid = (int)((tiles + tiles_per_cta - 1) / tiles_per_cta);
  if (one_lane) return pred3 ? launch_loop<true>(a, grid, stream) : launch_loop<false>(a, grid, stream);
  return pred3 ? launch_loop2<true>(a, grid, stream) : launch_loop2<false>(a, grid, stream);
}
