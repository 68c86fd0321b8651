// pipes.cu -- measured FP32 / FFMA2 / MUFU issue rates on this B200 (the denominators of the series
// kernel's roofline).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Each kernel runs `iters` iterations of an unrolled body of independent chains on a full grid
// (8 CTAs x 256 threads per SM); rates are lane-operations per second from CUDA-event time, and per
// clock per SM assuming the SM clock printed by nvidia-smi under load (passed as argv[1], MHz).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void fma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rd, {%0,%1};\n"
               "fma.rn.f32x2 rd, ra, rb, rd;\n mov.b64 {%0,%1}, rd;}\n"
               : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ void mul2(float& dx, float& dy, float bx, float by) {
  asm volatile("{.reg .b64 rb, rd;\n mov.b64 rb, {%2,%3};\n mov.b64 rd, {%0,%1};\n"
               "mul.rn.f32x2 rd, rd, rb;\n mov.b64 {%0,%1}, rd;}\n"
               : "+f"(dx), "+f"(dy) : "f"(bx), "f"(by));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__constant__ float ctab[64];

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = 1.0f + i * 1e-3f; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 FFMA d = a*b+c, three distinct registers
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], b[i], b[(i + 1) & 7]); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], a[i], a[(i + 3) & 7]); }
    } else if (MODE == 1) {  // 8 FFMA2 = 16 lane-FMAs
#pragma unroll
      for (int i = 0; i < 8; i += 2) fma2(a[i], a[i + 1], b[i], b[i + 1], a[(i + 2) & 7], a[(i + 3) & 7]);
#pragma unroll
      for (int i = 0; i < 8; i += 2) fma2(b[i], b[i + 1], a[i], a[i + 1], b[(i + 2) & 7], b[(i + 3) & 7]);
    } else if (MODE == 2) {  // 8 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]);
    } else if (MODE == 3) {  // 13 FFMA + 1 MUFU (the scalar series loop's mix)
#pragma unroll
      for (int i = 0; i < 7; ++i) { a[i] = fmaf(a[i], b[i], b[(i + 1) & 7]); }
#pragma unroll
      for (int i = 0; i < 6; ++i) { b[i] = fmaf(b[i], a[i], a[(i + 3) & 7]); }
      a[7] = ex2(b[7]); b[7] = a[7] * 0.999f;
    } else if (MODE == 5) {  // 16 FFMA with two immediates (one register read)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], 0.999f, 1e-3f); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], 1.001f, -1e-3f); }
    } else if (MODE == 6) {  // 16 FFMA d = d*b + imm (two register reads)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], b[i], 1e-3f); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], a[(i + 3) & 7], -1e-3f); }
    } else if (MODE == 7) {  // 16 FMUL d = d*b (two register reads)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = a[i] * b[i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = b[i] * a[(i + 3) & 7]; }
    } else if (MODE == 8) {  // 16 FADD d = d+b (two register reads)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = a[i] + b[i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = b[i] - a[(i + 3) & 7]; }
    } else if (MODE == 9) {  // 16 FFMA d = d*c[const] + b (two register reads + constant bank)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], ctab[i], b[i]); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], ctab[8 + i], a[(i + 3) & 7]); }
    } else if (MODE == 10) {  // 16 FFMA d = a*a + d style (repeated register: d = d*d + b)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], a[i], b[i]); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], b[i], a[(i + 3) & 7]); }
    } else if (MODE == 11) {  // 8 FMUL2 (two 64-bit register reads each)
#pragma unroll
      for (int i = 0; i < 8; i += 2) mul2(a[i], a[i + 1], b[i], b[i + 1]);
#pragma unroll
      for (int i = 0; i < 8; i += 2) mul2(b[i], b[i + 1], a[(i + 2) & 7], a[(i + 3) & 7]);
    } else if (MODE == 12) {  // 16 FFMA sharing one multiplicand in the same slot (reuse cache): d = d*s + b
      const float s = b[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) { a[i] = fmaf(s, a[i], b[i]); }
#pragma unroll
      for (int i = 0; i < 7; ++i) { b[i] = fmaf(s, b[i], a[(i + 3) & 7]); }
      a[7] = fmaf(s, a[7], b[0]); b[7] = fmaf(s, a[7], b[1]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter, double mufu_per_iter, int sms, double mhz) {
  float* out;
  const int blocks = sms * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  k<MODE><<<blocks, 256>>>(out, 2000);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double threads = (double)blocks * 256;
  const double fp = lane_ops_per_iter * iters * threads / (best * 1e-3), mu = mufu_per_iter * iters * threads / (best * 1e-3);
  printf("{\"test\": \"%s\", \"ms\": %.3f, \"fp32_lane_instr_per_s\": %.4g, \"mufu_per_s\": %.4g, \"fp32_lanes_per_clk_per_sm\": %.1f, \"mufu_per_clk_per_sm\": %.2f}\n",
         name, best, fp, mu, fp / (sms * mhz * 1e6), mu / (sms * mhz * 1e6));
  cudaFree(out);
}

int main(int argc, char** argv) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-3f * i;
  cudaMemcpyToSymbol(ctab, h, sizeof(h));
  printf("{\"device\": \"%s\", \"sms\": %d, \"assumed_sm_mhz\": %.0f}\n", p.name, sms, mhz);
  run<0>("ffma_rrr", 16, 0, sms, mhz);
  run<6>("ffma_rr_imm", 16, 0, sms, mhz);
  run<5>("ffma_r_imm_imm", 16, 0, sms, mhz);
  run<9>("ffma_rr_const", 16, 0, sms, mhz);
  run<10>("ffma_repeated_reg", 16, 0, sms, mhz);
  run<12>("ffma_shared_multiplicand", 16, 0, sms, mhz);
  run<7>("fmul_rr", 16, 0, sms, mhz);
  run<8>("fadd_rr", 16, 0, sms, mhz);
  run<1>("ffma2_rrr", 16, 0, sms, mhz);
  run<11>("fmul2_rr", 16, 0, sms, mhz);
  run<2>("mufu_ex2", 0, 8, sms, mhz);
  run<3>("mix_13ffma_1mufu", 13, 1, sms, mhz);
  return 0;
}
