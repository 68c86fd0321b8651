// so3d_denoiser.cu -- the RotPredict denoiser (so3_train.py:11-49, bingham_train.py:9-47: 65-wide, 4 x (Linear + SiLU)
// + Linear -> skew vector) fused with the reverse step of SO3Diffusion.p_sample (diffusion.py:291-326) in ONE kernel
// per step.  Stock PyTorch runs this step as ~20 launches that stream ~2.6 KB of activations per
// particle through HBM; here a particle costs 36 B in + 36 B out and the activations never leave the SM.
//
// Execution model (B200, tcgen05 + TMEM):
//   * persistent grid, one CTA of 18 warps per SM; the five weight matrices (tf32 hi/lo split, canonical K-major
//     no-swizzle UMMA layout, 158 KB) are staged ONCE per CTA in shared memory by bulk copies (TMA engine);
//   * warps 0..15 are two groups of 8 epilogue warps; a group owns a tile of 128 particles = the 128 TMEM lanes and
//     works on it with TWO threads per particle (each takes half of the 65 columns), so four warps per scheduler
//     hide the MUFU / tensor-memory latencies.  While one group's MMAs run on the tensor core the other group does
//     its SiLU epilogue on the FP32/XU pipes;
//   * warps 16, 17 are the MMA issuers of the two groups: they wait on an mbarrier for "A operand written" (one
//     arrival per epilogue warp), issue the layer under elect.sync -- descriptors stay in uniform registers, the 27
//     UTCHMMA of a layer go out back to back -- and tcgen05.commit to the group's "accumulator ready" mbarrier;
//   * every layer is D[128 x N] = A[128 x K] * W^T on the 5th-gen tensor core: `tcgen05.mma.kind::tf32`, A operand
//     read from TENSOR MEMORY (the previous layer's activations, written there by the epilogue with tcgen05.st --
//     no shared-memory round trip, no layout shuffle), B = weights from shared memory, accumulator in TMEM;
//   * fp32 accuracy from the tf32 tensor core by a 3-term split  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
//     (hi = x rounded to nearest at 10 mantissa bits, lo = exact remainder; dropped terms are O(2^-22)): three
//     accumulating MMAs per K step;
//   * bias = one more K column (activation column 65 is the constant 1); the 56-wide sinusoidal time embedding
//     (models.py:13-25) of the step's shared t is folded into the first layer's bias column (c1_table, built once
//     per weight set by the caller), so layer 1 is a K = 16 product of the 9 matrix entries;
//   * epilogue per layer: tcgen05.ld the thread's 32 / 33 accumulators, SiLU (MUFU.EX2 + MUFU.RCP), split,
//     tcgen05.st as the next A operand; after layer 5 the 3 outputs feed the quaternion reverse step
//     (so3d_math.cuh: p_mean_quat) in the particle's first thread, while its second thread has already drawn the
//     step's noise rotation (Philox, shared-row inverse CDF) and parked it in shared memory.
// Per 128 particles: 3 x (3 x 9) + 3 x 2 MMAs of 128 x 80 x 8 and 3 x 9 of 128 x 16 x 8 (~2.1 k tensor-pipe clocks
// measured) and 4 x 65 x 2 MUFU per particle (4.2 k XU-pipe clocks): the kernel is bound by the XU pipe (62 % busy,
// profiles/r01i_denoiser_full.md), then by instruction issue (51 %).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/so3d.h"
#include "so3d_cdf_smem.cuh"
#include "so3d_common.cuh"
#include "so3d_math.cuh"
#include "so3d_tma.cuh"

using namespace so3d;

namespace {

// ---- network geometry (RotPredict(d_model=65, out_type="skewvec"), so3_train.py:11-37) ------------------------------
constexpr int kD = SO3D_ROTPREDICT_D;        // 65: width of every hidden layer
constexpr int kIn = 9;                       // flattened rotation matrix
constexpr int kOut = 3;                      // skew vector
constexpr int kKPad = 72;                    // 65 activations + 1 bias column, padded to 9 K-steps of 8
constexpr int kNPad = 80;                    // 65 outputs padded to a legal UMMA N (multiple of 16 at M = 128)
constexpr int kK1 = 16;                      // layer 1: 9 inputs + 1 bias/time column, padded to 2 K-steps
constexpr int kN5 = 16;                      // layer 5: 3 outputs padded to the smallest UMMA N
constexpr int kM = 128;                      // particles per tile = TMEM lanes

// packed weights (floats): per layer [hi | lo], each in canonical K-major order (k/4, n, k%4)
constexpr int kL1Floats = kK1 * kNPad;       // 1280
constexpr int kLhFloats = kKPad * kNPad;     // 5760
constexpr int kL5Floats = kKPad * kN5;       // 1152
constexpr int kOffL1 = 0;
constexpr int kOffLh = kOffL1 + 2 * kL1Floats;                 // layers 2..4
constexpr int kOffL5 = kOffLh + 3 * 2 * kLhFloats;
constexpr int kBlobFloats = kOffL5 + 2 * kL5Floats;            // 39424 floats = 157 696 B
static_assert(kBlobFloats == SO3D_ROTPREDICT_BLOB_FLOATS, "so3d.h out of sync");

// TMEM columns of a group: accumulator, A hi, A lo
constexpr uint32_t kColD = 0, kColAhi = 96, kColAlo = 176, kColsPerGroup = 256;

constexpr size_t kSmemBytes = sizeof(float) * (size_t)(kBlobFloats + ((kTabCdfFloats + 3) & ~3)) + 2 * 2 * 128 * 16 /* noise quaternions */ + 64;  // 5 barriers, TMEM base, 2 counters

__device__ __forceinline__ int canon_index(int n, int k, int N) { return (k >> 2) * (N * 4) + n * 4 + (k & 3); }

// ---- tcgen05 wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle; element (n, k) at (k/4) * lbo + (n/8) * sbo + (n%8) * 16 + (k%4) * 4 bytes
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A and B tf32, both K-major, dense, M x N
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define SO3D_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define SO3D_W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
// 16 consecutive columns of this thread's lane -> v[o .. o+15]
#define SO3D_TMEM_LD16(taddr, v, o)                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"    \
               : SO3D_R8(v, o), SO3D_R8(v, o + 8)                                                                          \
               : "r"(taddr)                                                                                                \
               : "memory")
#define SO3D_TMEM_LD1(taddr, v, o) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[o]) : "r"(taddr) : "memory")
#define SO3D_TMEM_LD4(taddr, v, o) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]) : "r"(taddr) : "memory")
#define SO3D_TMEM_ST8(taddr, v, o) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), SO3D_W8(v, o) : "memory")

// tf32 split with round-to-nearest: hi = x rounded to 10 mantissa bits (low 13 bits zero), lo = x - hi exactly, so
// |lo| <= 2^-11 |x| with either sign; the tensor core then truncates lo to 10 bits (an UNBIASED 2^-21 |x| error --
// truncating hi instead would leave lo >= 0 and bias every product low by ~2^-22).
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

// x * sigmoid(x) (torch.nn.SiLU): one MUFU.EX2 and one MUFU.RCP.  (Measured: sharing one MUFU.RCP between two
// activations -- r = 1/(a0 a1), 1/a0 = r a1 -- is slower, 3.97 vs 3.37 ms per 2^24-particle step: the extra live values
// spill at the 96-register cap of this 18-warp CTA, profiles/r02a_probe_denoiser.log.)
__device__ __forceinline__ float silu(float x) {
  const float e = fast_ex2(-1.4426950408889634f * x);
  return x * rcp_approx(1.0f + e);
}

__device__ __forceinline__ void mbar_wait_bounded_addr(uint32_t addr, uint32_t parity);
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) { mbar_wait_bounded_addr(smem_u32(bar), parity); }
// the barrier's shared-memory address as an opaque register value: ptxas otherwise folds it into another pointer plus a
// NEGATIVE immediate ([R44 + -0x68] in the multi-step kernel), an addressing form compute-sanitizer's synccheck does not
// resolve -- it then reports the (initialised) barrier as "Missing init" (profiles/r03c_sanitizer_synccheck_*.log)
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }  // (a PTX mov is folded away again)
__device__ __forceinline__ void mbar_wait_bounded_addr(uint32_t addr, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) __trap();  // a tensor-core op that never completes must not hang the device
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct DenoiseArgs {
  const float* x_t;
  const float* blob;
  const float* c1_table;
  const int64_t* t;
  const float* recip;
  const float* recipm1;
  const float* coef1;
  const float* coef2;
  int64_t T;
  const float* post_cdf;
  const float* loc;
  uint64_t seed, rng_offset, row_offset;
  const uint64_t* seed_dev;  // non-null: the seed lives in device memory (CUDA-graph replays draw fresh noise)
  int64_t t_hi, t_lo;        // kLoop: the kernel runs the steps t_hi, t_hi - 1, ..., t_lo itself (rng_offset = step)
  float* out;
  float* pred_out;
  int64_t n;
};

// issue the MMAs of one layer for a group (one thread): D = A_hi B_hi + A_lo B_hi + A_hi B_lo over `ksteps` K-steps
__device__ __forceinline__ void issue_layer(uint32_t tmem_group, uint32_t b_hi_addr, uint32_t b_lo_addr, int N, int ksteps, uint64_t* bar) {
  const uint32_t lbo = (uint32_t)N * 16u, sbo = 128u;
  const uint32_t idesc = instr_desc(kM, N);
  tc_fence_after();
  for (int s = 0; s < ksteps; ++s) {
    const uint64_t bh = smem_desc(b_hi_addr + (uint32_t)s * 2u * lbo, lbo, sbo);
    const uint64_t bl = smem_desc(b_lo_addr + (uint32_t)s * 2u * lbo, lbo, sbo);
    const uint32_t ah = tmem_group + kColAhi + (uint32_t)s * 8u, al = tmem_group + kColAlo + (uint32_t)s * 8u;
    mma_tf32_ts(tmem_group + kColD, ah, bh, idesc, s > 0);
    mma_tf32_ts(tmem_group + kColD, al, bh, idesc, 1u);
    mma_tf32_ts(tmem_group + kColD, ah, bl, idesc, 1u);
  }
  tc_commit(bar);
}

// SiLU epilogue of one hidden layer for the thread's particle (TMEM lane) and its half of the columns:
// H = 0: accumulator columns 0..31 -> A columns 0..31;  H = 1: columns 32..64 -> A columns 32..71 (column 65 is the
// constant 1 that carries the next layer's bias, 66..71 are zero padding).
template <int H>
__device__ __forceinline__ void silu_epilogue(uint32_t tmem_lane) {
  constexpr int kC0 = H ? 4 : 0, kChunks = H ? 5 : 4;
  uint32_t acc[40];
  SO3D_TMEM_LD16(tmem_lane + kColD + kC0 * 8, acc, 0);
  SO3D_TMEM_LD16(tmem_lane + kColD + kC0 * 8 + 16, acc, 16);
  if (H) SO3D_TMEM_LD1(tmem_lane + kColD + 64, acc, 32);
  tc_wait_ld();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int col = (kC0 + c) * 8 + k;
      float v;
      if (col < kD)
        v = silu(__uint_as_float(acc[c * 8 + k]));
      else
        v = col == kD ? 1.0f : 0.0f;
      hi[k] = tf32_hi(v);
      lo[k] = tf32_lo(v, hi[k]);
    }
    SO3D_TMEM_ST8(tmem_lane + kColAhi + (kC0 + c) * 8, hi, 0);
    SO3D_TMEM_ST8(tmem_lane + kColAlo + (kC0 + c) * 8, lo, 0);
  }
  tc_wait_st();
}

// Roles: warps 0..15 are two groups of 8 epilogue warps (group g = warp / 8 owns one 128-particle tile at a time);
// inside a group warp w handles TMEM lanes 32 (w % 4) .. +31 (the hardware's lane-quarter rule) and column half
// h = (w / 4) % 2 -- two threads per particle.  Warps 16 and 17 issue the MMAs of group 0 and 1.  Hand-offs are
// mbarriers: bar_a[g] "A operand written" (8 warp arrivals) -> issuer -> tcgen05.commit -> bar_d[g] "accumulator ready".
// (SO3D_DENOISER_ISSUER_WARPS=0 builds the variant without issuer warps: the LAST of the group's 8 warps to finish a
// layer's operand -- an acq_rel counter in shared memory tells it -- issues the MMAs itself; see the knob above.)
#ifndef SO3D_DENOISER_ISSUER_WARPS
#define SO3D_DENOISER_ISSUER_WARPS 1  // 1 (shipped): two dedicated issuer warps (18 warps: 96 registers per thread, 12-40 B of spills);
#endif                                // 0: the last epilogue warp to finish a layer's operand issues the next layer's MMAs itself (16 warps,
                                      //    128 registers) -- measured SLOWER, 3.93 vs 3.33 ms per 2^24-particle step
                                      //    (profiles/r03q_denoiser_issuer_negative.jsonl): the 27 MMA issues per layer then sit on an epilogue
                                      //    warp's critical path instead of overlapping the other group's epilogue
constexpr int kEpiWarps = 16, kThreads = (kEpiWarps + (SO3D_DENOISER_ISSUER_WARPS ? 2 : 0)) * 32;

// kLoop = false: one reverse step (t = a.t[0]).  kLoop = true: the WHOLE reverse process t_hi .. t_lo in one launch --
// particles are independent and every particle is always handled by the same thread of the same CTA, so the steps
// need no grid-wide synchronisation: the CTA keeps the weights in shared memory, restages the step's CDF row and
// time-embedding column between steps (two CTA barriers), reads its rows back from `out` (written by the same thread
// one step earlier) and draws the step's noise with rng_offset = step, i.e. exactly the launches of
// diffusion.py:328-337 with so3d_rotpredict_p_sample_f32(..., rng_offset = t) per step, minus 999 launches, 999 weight
// loads and the HBM round trip of the state between steps.
template <bool kLoop>
__global__ void __launch_bounds__(kThreads, 1) rotpredict_p_sample_kernel(const DenoiseArgs a) {
  extern __shared__ float4 smem4[];
  float* s_blob = reinterpret_cast<float*>(smem4);
  float* s_tab = s_blob + kBlobFloats;
  float4* s_noise = reinterpret_cast<float4*>(s_tab + ((kTabCdfFloats + 3) & ~3));   // [group][parity][particle]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_noise + 2 * 2 * kM);  // [0] weights, [1..2] bar_d, [3..4] bar_a
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 5);
  uint32_t* s_cnt = s_tmem + 2;  // [group]: warps that have finished the current layer's A operand (runs on, mod 8)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int64_t ti = kLoop ? a.t_hi : a.t[0];
  ti = ti < 0 ? 0 : (ti >= a.T ? a.T - 1 : ti);
  bool noisy = a.post_cdf && ti != 0 && a.out;

  if (warp == 0) tmem_alloc(s_tmem, 512);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 8);
    mbar_init(&bars[4], 8);
    s_cnt[0] = 0;
    s_cnt[1] = 0;
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {  // weights: global -> shared through the TMA engine, 5 bulk copies
    mbar_expect_tx(&bars[0], (uint32_t)(kBlobFloats * sizeof(float)));
    constexpr int kPiece = 8192;  // floats (32 KB)
    for (int off = 0; off < kBlobFloats; off += kPiece) {
      const int cnt = (kBlobFloats - off) < kPiece ? (kBlobFloats - off) : kPiece;
      bulk_load(s_blob + off, a.blob + off, (uint32_t)(cnt * sizeof(float)), &bars[0]);
    }
  }
  mbar_wait(&bars[0], 0);
  const uint32_t tmem_base = *s_tmem;
  const uint32_t blob_addr = smem_u32(s_blob);
  const int64_t tiles = (a.n + kM - 1) / kM;
  uint32_t ph = 0;   // mbarrier phase of this warp's hand-off barrier: runs on across the steps
  int par = 0;       // parity of the noise staging buffer: likewise
  const int64_t step_lo = kLoop ? (a.t_lo < 0 ? 0 : a.t_lo) : ti;
  int64_t step = ti;
  do {  // kLoop = false: the body runs once and the loop folds away
  if (kLoop) {
    if (step != ti) {  // every warp is done with the previous step's tables, bias column and accumulators
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    noisy = a.post_cdf && step != 0 && a.out;
  }
  const int64_t ts = step;  // the step index this pass works on
  const float* __restrict__ src = (kLoop && step != ti) ? a.out : a.x_t;
  const uint64_t step_offset = kLoop ? (uint64_t)step : a.rng_offset;
  if (noisy) stage_cdf(s_tab, a.post_cdf + ts * kCdf, a.loc);  // posterior CDF row of this step + guide
  // time embedding of this step folded into layer 1's bias column (k = 9): generic-proxy writes, ordered before the MMAs
  if (tid < kD) {
    const float c = __ldg(a.c1_table + ts * kD + tid);
    const uint32_t hi = tf32_hi(c);
    s_blob[kOffL1 + canon_index(tid, kIn, kNPad)] = __uint_as_float(hi);
    s_blob[kOffL1 + kL1Floats + canon_index(tid, kIn, kNPad)] = __uint_as_float(tf32_lo(c, hi));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (SO3D_DENOISER_ISSUER_WARPS && warp >= kEpiWarps) {
    // ---- MMA issuer of group g: wait for the A operand, issue the layer, commit to the accumulator barrier ----
    const int g = warp - kEpiWarps;
    const uint32_t tmem_group = tmem_base + (uint32_t)g * kColsPerGroup;
    const uint32_t bar_a_addr = opaque_u32(smem_u32(&bars[3 + g]));
    for (int64_t tile = (int64_t)blockIdx.x * 2 + g; tile < tiles; tile += (int64_t)gridDim.x * 2) {
      {
#pragma unroll 1
        for (int layer = 1; layer <= 5; ++layer) {
          mbar_wait_bounded_addr(bar_a_addr, ph);
          ph ^= 1;
          if (!elect_one()) continue;
          if (layer == 1) {
            issue_layer(tmem_group, blob_addr + kOffL1 * 4, blob_addr + (kOffL1 + kL1Floats) * 4, kNPad, kK1 / 8, &bars[1 + g]);
          } else if (layer < 5) {
            const uint32_t b = blob_addr + (uint32_t)(kOffLh + (layer - 2) * 2 * kLhFloats) * 4u;
            issue_layer(tmem_group, b, b + kLhFloats * 4u, kNPad, kKPad / 8, &bars[1 + g]);
          } else {
            const uint32_t b = blob_addr + (uint32_t)kOffL5 * 4u;
            issue_layer(tmem_group, b, b + kL5Floats * 4u, kN5, kKPad / 8, &bars[1 + g]);
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue warps ----
    const int g = warp >> 3, h = (warp >> 2) & 1, r = (warp & 3) * 32 + lane;
    const uint32_t tmem_group = tmem_base + (uint32_t)g * kColsPerGroup;
    const uint32_t tmem_lane = tmem_group + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 lanes of the group's columns
    uint64_t* bar_d = &bars[1 + g];
    uint64_t* bar_a = &bars[3 + g];
    // this warp's part of layer `layer`'s A operand is in tensor memory: tell the issuer -- or BE the issuer, if last
    auto operand_done = [&](int layer) {
      tc_fence_before();
      __syncwarp();
      if (SO3D_DENOISER_ISSUER_WARPS) {
        if (lane == 0) mbar_arrive(bar_a);
        return;
      }
      uint32_t old = 0;
      if (lane == 0) asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&s_cnt[g])) : "memory");
      old = __shfl_sync(0xffffffffu, old, 0);
      if ((old & 7u) != 7u) return;
      if (elect_one()) {  // (issue_layer starts with tcgen05.fence::after_thread_sync)
        if (layer == 1) {
          issue_layer(tmem_group, blob_addr + kOffL1 * 4, blob_addr + (kOffL1 + kL1Floats) * 4, kNPad, kK1 / 8, bar_d);
        } else if (layer < 5) {
          const uint32_t b = blob_addr + (uint32_t)(kOffLh + (layer - 2) * 2 * kLhFloats) * 4u;
          issue_layer(tmem_group, b, b + kLhFloats * 4u, kNPad, kKPad / 8, bar_d);
        } else {
          const uint32_t b = blob_addr + (uint32_t)kOffL5 * 4u;
          issue_layer(tmem_group, b, b + kL5Floats * 4u, kN5, kKPad / 8, bar_d);
        }
      }
      __syncwarp();
    };
    const float k_recip = __ldg(a.recip + ts), k_recipm1 = __ldg(a.recipm1 + ts);
    const float k_c1 = __ldg(a.coef1 + ts), k_c2 = __ldg(a.coef2 + ts);

    const int64_t tile0 = (int64_t)blockIdx.x * 2 + g, tstride = (int64_t)gridDim.x * 2;
    Mat3 x = identity(), xn = identity();
    if (h == 0 && tile0 < tiles && tile0 * kM + r < a.n) {
#pragma unroll
      for (int k = 0; k < 9; ++k) x.m[k] = __ldcs(src + (tile0 * kM + r) * 9 + k);
    }
    for (int64_t tile = tile0; tile < tiles; tile += tstride, par ^= 1) {
      const int64_t i = tile * kM + r;
      const bool live = i < a.n;
      if (h == 0) {
        // layer 1 operand: [x(9), 1, 0 x 6]
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float v = k < 9 ? x.m[k] : (k == 9 ? 1.0f : 0.0f);
          hi[k] = tf32_hi(v);
          lo[k] = tf32_lo(v, hi[k]);
        }
        SO3D_TMEM_ST8(tmem_lane + kColAhi, hi, 0);
        SO3D_TMEM_ST8(tmem_lane + kColAhi + 8, hi, 8);
        SO3D_TMEM_ST8(tmem_lane + kColAlo, lo, 0);
        SO3D_TMEM_ST8(tmem_lane + kColAlo + 8, lo, 8);
        tc_wait_st();
      }
      operand_done(1);
      if (h == 0) {
        // prefetch the next tile's rotation while this tile runs through the network
        const int64_t in = (tile + tstride) * kM + r;
        xn = identity();
        if (tile + tstride < tiles && in < a.n) {
#pragma unroll
          for (int k = 0; k < 9; ++k) xn.m[k] = __ldcs(src + in * 9 + k);
        }
      } else if (noisy && live) {
        // the step's noise rotation does not depend on the network: drawn by the second thread of the particle
        const uint64_t sd = a.seed_dev ? __ldg(reinterpret_cast<const unsigned long long*>(a.seed_dev)) : a.seed;
        const NoiseDraw d = draw_axis_u(sd, a.row_offset + (uint64_t)i, step_offset);
        const Quat qn = quat_axis_angle(d.axis, shared_row_angle(s_tab, d.u));
        s_noise[(g * 2 + par) * kM + r] = make_float4(qn.w, qn.x, qn.y, qn.z);
      }
#pragma unroll 1
      for (int layer = 2; layer <= 5; ++layer) {
        mbar_wait_bounded(bar_d, ph);
        ph ^= 1;
        tc_fence_after();
        if (h == 0)
          silu_epilogue<0>(tmem_lane);
        else
          silu_epilogue<1>(tmem_lane);
        operand_done(layer);
      }
      // ---- output of layer 5 = predicted skew vector; fused reverse step (diffusion.py:291-326) ----
      mbar_wait_bounded(bar_d, ph);
      ph ^= 1;
      tc_fence_after();
      uint32_t pr[4] = {0u, 0u, 0u, 0u};
      if (h == 0) {
        SO3D_TMEM_LD4(tmem_lane + kColD, pr, 0);
        tc_wait_ld();
      }
      tc_fence_before();  // the next tile's tcgen05.st / MMAs of this group are ordered after these loads by bar_a
      if (noisy) asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory");  // the group's noise quaternions are in shared memory
      if (h == 0) {
        const Vec3 pred{__uint_as_float(pr[0]), __uint_as_float(pr[1]), __uint_as_float(pr[2])};
        if (live) {
          if (a.pred_out) {
            a.pred_out[i * 3 + 0] = pred.x;
            a.pred_out[i * 3 + 1] = pred.y;
            a.pred_out[i * 3 + 2] = pred.z;
          }
          if (a.out) {
            Quat qh;
            Quat qm = p_mean_quat(x, pred, k_recip, k_recipm1, k_c1, k_c2, &qh);
            if (noisy) {
              const float4 qn = s_noise[(g * 2 + par) * kM + r];
              qm = qmul(qm, Quat{qn.x, qn.y, qn.z, qn.w});
            }
            const Mat3 o = quat_to_mat_unit(qm);
#pragma unroll
            for (int k = 0; k < 9; ++k) __stcs(a.out + i * 9 + k, o.m[k]);
          }
        }
        x = xn;
      }
    }
  }
  } while (kLoop && --step >= step_lo);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- weight packing: nn.Linear weights (out x in, row-major) -> tf32 hi/lo blob in the UMMA layout ------------------
struct PackArgs {
  const float* w[5];
  const float* b[5];
  float* blob;
};

__device__ __forceinline__ void put_split(float* blob, int hi_off, int lo_off, int idx, float v) {
  const uint32_t hi = tf32_hi(v);
  blob[hi_off + idx] = __uint_as_float(hi);
  blob[lo_off + idx] = __uint_as_float(tf32_hi(v - __uint_as_float(hi)));  // lo rounded (not truncated) to tf32 as well
}

__global__ void __launch_bounds__(256) rotpredict_pack_kernel(const PackArgs p) {
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kBlobFloats / 2; e += stride) {
    // e enumerates the (layer, n, k) slots of the hi halves; the lo half is written alongside
    int q = e;
    if (q < kL1Floats) {  // layer 1: columns 0..8 = W1[:, :9]; column 9 (time/bias) is filled per step by the kernel
      const int n = q / kK1, k = q % kK1;
      const float v = (n < kD && k < kIn) ? p.w[0][n * kD + k] : 0.0f;
      put_split(p.blob, kOffL1, kOffL1 + kL1Floats, canon_index(n, k, kNPad), v);
      continue;
    }
    q -= kL1Floats;
    if (q < 3 * kLhFloats) {  // layers 2..4: columns 0..64 = W, column 65 = bias
      const int l = q / kLhFloats, qq = q % kLhFloats;
      const int n = qq / kKPad, k = qq % kKPad;
      float v = 0.0f;
      if (n < kD) v = k < kD ? p.w[1 + l][n * kD + k] : (k == kD ? p.b[1 + l][n] : 0.0f);
      const int base = kOffLh + l * 2 * kLhFloats;
      put_split(p.blob, base, base + kLhFloats, canon_index(n, k, kNPad), v);
      continue;
    }
    q -= 3 * kLhFloats;
    {  // layer 5: 3 outputs
      const int n = q / kKPad, k = q % kKPad;
      float v = 0.0f;
      if (n < kOut) v = k < kD ? p.w[4][n * kD + k] : (k == kD ? p.b[4][n] : 0.0f);
      put_split(p.blob, kOffL5, kOffL5 + kL5Floats, canon_index(n, k, kN5), v);
    }
  }
}

}  // namespace

#define SO3D_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) return so3d_host::fail(SO3D_EINVAL, msg); \
  } while (0)

extern "C" {

int so3d_rotpredict_pack_f32(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                             const float* w4, const float* b4, const float* w5, const float* b5, float* blob, void* stream) {
  SO3D_REQUIRE(w1 && b1 && w2 && b2 && w3 && b3 && w4 && b4 && w5 && b5 && blob, "so3d_rotpredict_pack_f32: null pointer");
  PackArgs p;
  p.w[0] = w1; p.w[1] = w2; p.w[2] = w3; p.w[3] = w4; p.w[4] = w5;
  p.b[0] = b1; p.b[1] = b2; p.b[2] = b3; p.b[3] = b4; p.b[4] = b5;
  p.blob = blob;
  rotpredict_pack_kernel<<<(kBlobFloats / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
  return so3d_host::check_launch("so3d_rotpredict_pack_f32");
}

static int rotpredict_p_sample_impl(const float* x_t, const float* blob, const float* c1_table, const int64_t* t, const float* recip,
                                    const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                                    const float* loc, uint64_t seed, const uint64_t* seed_dev, uint64_t rng_offset, uint64_t row_offset,
                                    float* out, float* pred_out, int64_t n, void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(x_t && blob && c1_table && t && recip && recipm1 && coef1 && coef2, "so3d_rotpredict_p_sample_f32: null pointer");
  SO3D_REQUIRE(out || pred_out, "so3d_rotpredict_p_sample_f32: no output requested");
  SO3D_REQUIRE(T > 0, "so3d_rotpredict_p_sample_f32: T must be positive");
  SO3D_REQUIRE(!post_cdf || loc, "so3d_rotpredict_p_sample_f32: loc required with post_cdf");
  SO3D_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15u) == 0, "so3d_rotpredict_p_sample_f32: blob must be 16-byte aligned");
  static bool configured_dev[64] = {};  // the dynamic shared-memory limit is a per-device kernel attribute
  bool& configured = configured_dev[so3d_host::current_device()];
  if (!configured) {
    if (cudaFuncSetAttribute(rotpredict_p_sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
      return so3d_host::check_launch("so3d_rotpredict_p_sample_f32 (shared memory)");
    configured = true;
  }
  DenoiseArgs a;
  a.x_t = x_t; a.blob = blob; a.c1_table = c1_table; a.t = t; a.recip = recip; a.recipm1 = recipm1; a.coef1 = coef1; a.coef2 = coef2;
  a.T = T; a.post_cdf = post_cdf; a.loc = loc; a.seed = seed; a.seed_dev = seed_dev; a.rng_offset = rng_offset; a.row_offset = row_offset;
  a.out = out; a.pred_out = pred_out; a.n = n;
  const int64_t pairs = ((n + kM - 1) / kM + 1) / 2;
  const int sms = so3d_host::sm_count();
  const int grid = (int)(pairs < sms ? pairs : sms);
  a.t_hi = a.t_lo = 0;
  rotpredict_p_sample_kernel<false><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a);
  return so3d_host::check_launch("so3d_rotpredict_p_sample_f32");
}

int so3d_rotpredict_p_sample_f32(const float* x_t, const float* blob, const float* c1_table, const int64_t* t, const float* recip,
                                 const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                                 const float* loc, uint64_t seed, uint64_t rng_offset, uint64_t row_offset, float* out,
                                 float* pred_out, int64_t n, void* stream) {
  return rotpredict_p_sample_impl(x_t, blob, c1_table, t, recip, recipm1, coef1, coef2, T, post_cdf, loc, seed, nullptr, rng_offset, row_offset,
                                  out, pred_out, n, stream);
}

int so3d_rotpredict_p_sample_loop_f32(const float* x_t, const float* blob, const float* c1_table, int64_t t_hi, int64_t t_lo, const float* recip,
                                      const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                                      const float* loc, uint64_t seed, const uint64_t* seed_dev, uint64_t row_offset, float* out, int64_t n,
                                      void* stream) {
  SO3D_REQUIRE(n >= 0, "negative n");
  if (n == 0) return 0;
  SO3D_REQUIRE(x_t && blob && c1_table && recip && recipm1 && coef1 && coef2 && post_cdf && loc && out, "so3d_rotpredict_p_sample_loop_f32: null pointer");
  SO3D_REQUIRE(T > 0 && t_lo >= 0 && t_hi >= t_lo && t_hi < T, "so3d_rotpredict_p_sample_loop_f32: need 0 <= t_lo <= t_hi < T");
  SO3D_REQUIRE(out != x_t, "so3d_rotpredict_p_sample_loop_f32: out must not alias x_t");
  SO3D_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15u) == 0, "so3d_rotpredict_p_sample_loop_f32: blob must be 16-byte aligned");
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[so3d_host::current_device()];
  if (!configured) {
    if (cudaFuncSetAttribute(rotpredict_p_sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
      return so3d_host::check_launch("so3d_rotpredict_p_sample_loop_f32 (shared memory)");
    configured = true;
  }
  DenoiseArgs a;
  a.x_t = x_t; a.blob = blob; a.c1_table = c1_table; a.t = nullptr; a.recip = recip; a.recipm1 = recipm1; a.coef1 = coef1; a.coef2 = coef2;
  a.T = T; a.post_cdf = post_cdf; a.loc = loc; a.seed = seed; a.seed_dev = seed_dev; a.rng_offset = 0; a.row_offset = row_offset;
  a.t_hi = t_hi; a.t_lo = t_lo; a.out = out; a.pred_out = nullptr; a.n = n;
  const int64_t pairs = ((n + kM - 1) / kM + 1) / 2;
  const int sms = so3d_host::sm_count();
  const int grid = (int)(pairs < sms ? pairs : sms);
  rotpredict_p_sample_kernel<true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a);
  return so3d_host::check_launch("so3d_rotpredict_p_sample_loop_f32");
}

int so3d_rotpredict_p_sample_dseed_f32(const float* x_t, const float* blob, const float* c1_table, const int64_t* t, const float* recip,
                                       const float* recipm1, const float* coef1, const float* coef2, int64_t T, const float* post_cdf,
                                       const float* loc, const uint64_t* seed_dev, uint64_t rng_offset, uint64_t row_offset, float* out,
                                       float* pred_out, int64_t n, void* stream) {
  SO3D_REQUIRE(seed_dev, "so3d_rotpredict_p_sample_dseed_f32: null seed pointer");
  return rotpredict_p_sample_impl(x_t, blob, c1_table, t, recip, recipm1, coef1, coef2, T, post_cdf, loc, 0, seed_dev, rng_offset, row_offset,
                                  out, pred_out, n, stream);
}

}  // extern "C"
