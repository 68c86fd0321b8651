// series_variants.cu -- A/B harness for the inner loop of the IGSO(3) series kernel (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o series_variants series_variants.cu
//   ./series_variants [log2_rows]
// Every variant evaluates log f and d log f / dw of the L = 2000 series on the E-set (per-row omega, eps), is
// timed with CUDA events and checked against an fp64 host series on a sample.  V0 is the round-1a loop
// (rotation recurrence + 32-term anchors, kept verbatim below), V1 the shipped loop of so3d_math.cuh.
// Intermediate variants that were measured and dropped (constant-table operands on the V0 recurrence: 16.4
// clk/term; the same with two rows per thread packed as f32x2: 14.5 clk/term; V0: 18.2) are described in
// DESIGN.md section 4 -- FP32 throughput here is bounded by register-operand reads, which FFMA2 does not reduce.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../diffusion_extensions_b200/csrc/so3d_math.cuh"

using namespace so3d;

constexpr int L = 2000;

// ---- the round-1a loop (13 FP32 + 1 MUFU per term, every factor from the vector register file), kept here
// verbatim as the A/B baseline ---------------------------------------------------------------------------
namespace so3d {
constexpr int kAnchor = 32;

// sin/cos of (m * w) with the product carried exactly (hi + lo), first-order correction for lo.
SO3D_HD void sincos_mw_v0(float m, float w, float* s, float* c) {
  const float hi = m * w;
  const float lo = fmaf(m, w, -hi);
  float sh, ch;
  sincos_f(hi, &sh, &ch);
  *s = fmaf(lo, ch, sh);
  *c = fmaf(-lo, sh, ch);
}

struct SeriesAccV0 {
  float F, Fp;  // sum (l+1/2) e_l chi_l ,  sum (l+1/2) e_l D_l
};

struct SeriesStateV0 {
  float s, c, m;      // sin(m w), cos(m w), m as float
  float chi, D;       // running chi_m, D_m
  float F, Fp;        // accumulators
};

// `count` consecutive terms starting at the state's m: accumulate term m, then rotate to m + 1.
template <int kUnroll>
SO3D_HD void igso3_series_run_v0(SeriesStateV0& st, float sw, float cw, float cexp, int count) {
#pragma unroll kUnroll
  for (int i = 0; i < count; ++i) {
    st.chi = fmaf(2.0f, st.c, st.chi);
    st.D = fmaf(st.m, st.s, st.D);
    const float e = fast_ex2(fmaf(st.m, st.m, st.m) * cexp);
    const float p = e * (st.m + 0.5f);
    st.F = fmaf(p, st.chi, st.F);
    st.Fp = fmaf(p, st.D, st.Fp);
    const float t1 = st.c * sw, t2 = st.s * sw;
    const float sn = fmaf(st.s, cw, t1);
    st.c = fmaf(st.c, cw, -t2);
    st.s = sn;
    st.m += 1.0f;
  }
}

// Evaluate terms l = 0 .. L-1.  Returns F, Fp;  f = 2F, dlogf/dw = -2 Fp / F.
SO3D_HD SeriesAccV0 igso3_series_terms_v0(float w, float eps, int L) {
  const float cexp = -(eps * eps) * 1.4426950408889634f;
  float sw, cw;
  sincos_f(w, &sw, &cw);
  SeriesStateV0 st;
  st.s = sw; st.c = cw; st.m = 1.0f;
  st.chi = 1.0f; st.D = 0.0f;
  st.F = 0.5f; st.Fp = 0.0f;  // l = 0 term: (1/2) * e_0 * chi_0, D_0 = 0
  if (L > 1) igso3_series_run_v0<kAnchor>(st, sw, cw, cexp, (L < kAnchor ? L : kAnchor) - 1);
  for (int base = kAnchor; base < L; base += kAnchor) {
    sincos_mw_v0((float)base, w, &st.s, &st.c);
    st.m = (float)base;
    const int cnt = (L - base < kAnchor) ? (L - base) : kAnchor;
    if (cnt == kAnchor) igso3_series_run_v0<kAnchor>(st, sw, cw, cexp, kAnchor);
    else igso3_series_run_v0<1>(st, sw, cw, cexp, cnt);
  }
  return SeriesAccV0{st.F, st.Fp};
}

}  // namespace so3d

// ---- V0: round-1a loop ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) v0(const float* __restrict__ w, const float* __restrict__ eps, float* __restrict__ lf,
                                          float* __restrict__ g, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const SeriesAccV0 a = igso3_series_terms_v0(w[i], eps[i], L);
  lf[i] = logf(2.0f * a.F);
  g[i] = -2.0f * a.Fp / a.F;
}

// ---- V1: the shipped loop (so3d_math.cuh): Reinsch-Clenshaw recurrence, row-invariant factors as uniform operands ---
__global__ void __launch_bounds__(256) v1(const float* __restrict__ w, const float* __restrict__ eps, float* __restrict__ lf,
                                          float* __restrict__ g, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const SeriesAcc a = igso3_series_terms(w[i], eps[i], L);
  lf[i] = logf(2.0f * a.F);
  g[i] = a.dF / a.F;
}

static void truth(double w, double e, double* lf, double* g) {
  // character form in fp64: f = sum (2l+1) e_l chi_l, f' = -2 sum (2l+1) e_l D_l
  double chi = 1.0, D = 0.0, F = 0.5, Fp = 0.0;
  for (int m = 1; m < L; ++m) {
    chi += 2.0 * cos(m * w);
    D += m * sin(m * w);
    const double p = (m + 0.5) * exp(-(double)m * (m + 1) * e * e);
    F += p * chi;
    Fp += p * D;
  }
  *lf = log(2.0 * F);
  *g = -2.0 * Fp / F;
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 22;
  const int n = 1 << lg;
  std::vector<float> hw(n), he(n);
  srand(1234);
  for (int i = 0; i < n; ++i) {
    const double u1 = rand() / (RAND_MAX + 1.0), u2 = rand() / (RAND_MAX + 1.0);
    const double e = exp(log(6.4e-3) * (1.0 - u1));
    double om = e * sqrt(2.0) * 4.0 * u2;
    if (om > 3.0) om = 3.0;
    he[i] = (float)e;
    hw[i] = (float)om;
  }
  float *dw, *de, *dlf, *dg;
  cudaMalloc(&dw, n * 4); cudaMalloc(&de, n * 4); cudaMalloc(&dlf, n * 4); cudaMalloc(&dg, n * 4);
  cudaMemcpy(dw, hw.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(de, he.data(), n * 4, cudaMemcpyHostToDevice);
  const int ns = 4096;
  std::vector<double> tlf(ns), tg(ns);
  for (int k = 0; k < ns; ++k) truth((double)hw[k * (n / ns)], (double)he[k * (n / ns)], &tlf[k], &tg[k]);
  std::vector<float> hlf(n), hg(n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; ++v) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemset(dlf, 0, n * 4);
      cudaEventRecord(e0);
      if (v == 0) v0<<<n / 256, 256>>>(dw, de, dlf, dg, n);
      if (v == 1) v1<<<n / 256, 256>>>(dw, de, dlf, dg, n);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    cudaMemcpy(hlf.data(), dlf, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hg.data(), dg, n * 4, cudaMemcpyDeviceToHost);
    double ef = 0, eg = 0, ef_all = 0;
    for (int k = 0; k < ns; ++k) {
      const int i = k * (n / ns);
      const double kk = hw[i] / (sqrt(2.0) * he[i]);
      const double df = fabs(exp((double)hlf[i] - tlf[k]) - 1.0);
      const double dg_ = fabs((double)hg[i] - tg[k]) / fmax(fabs(tg[k]), 1e-30);
      if (df > ef_all) ef_all = df;
      if (kk <= 2.5) {
        if (df > ef) ef = df;
        if (hw[i] > 1e-4 && dg_ > eg) eg = dg_;
      }
    }
    printf("{\"variant\": %d, \"rows\": %d, \"ms\": %.3f, \"evals_per_s\": %.4g, \"clk_per_row_term_at_1965\": %.2f, \"max_rel_f_k<=2.5\": %.2e, \"max_rel_g_k<=2.5\": %.2e, \"max_rel_f_all\": %.2e}\n",
           v, n, best, n / (best * 1e-3), (best * 1e-3) * 1965e6 * 148 * 4 * 32 / ((double)n * L), ef, eg, ef_all);
  }
  return 0;
}
