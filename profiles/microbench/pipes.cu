// pipes.cu -- measured FP32 / FFMA2 / MUFU issue rates on this B200 (the denominators of the series
// kernel's roofline).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Each kernel runs `iters` iterations of an unrolled body of independent chains; the result is
// lane-operations per clock per SM, derived from CUDA-event time and the SM clock sampled by
// clock64() inside the kernel (so the figure does not depend on the boost state).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void fma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rd, {%0,%1};\n"
               "fma.rn.f32x2 rd, ra, rb, rd;\n mov.b64 {%0,%1}, rd;}\n"
               : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* clk) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = 1.0f + i * 1e-3f; }
  const float m = 0.999f, c = 1e-3f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 FFMA (3-register form)
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
      a[7] = ex2(b[7]); b[7] = a[7] * m;
    } else if (MODE == 4) {  // 13 FFMA2 + 2 MUFU (two rows per thread, packed)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) fma2(a[i], a[i + 1], b[i], b[i + 1], a[(i + 2) & 7], a[(i + 3) & 7]);
      }
      fma2(b[0], b[1], a[0], a[1], b[2], b[3]);
      b[6] = ex2(b[4]); b[7] = ex2(b[5]);
      b[4] = b[6] * m; b[5] = b[7] * m;
    } else if (MODE == 5) {  // 16 FFMA with a constant-bank/immediate operand (2-register form)
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], 0.999f, 1e-3f); }
#pragma unroll
      for (int i = 0; i < 8; ++i) { b[i] = fmaf(b[i], 1.001f, -1e-3f); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + m + c;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter, double mufu_per_iter, int sms) {
  float* out; long long* clk;
  const int blocks = sms * 8, iters = 20000;
  cudaMalloc(&out, blocks * 256 * sizeof(float)); cudaMalloc(&clk, 8);
  k<MODE><<<blocks, 256>>>(out, 100, clk);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, clk);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cycles; cudaMemcpy(&cycles, clk, 8, cudaMemcpyDeviceToHost);
  const double threads = (double)blocks * 256;
  const double mhz = cycles / (ms * 1e3);   // block 0's cycles ~ whole kernel (single wave)
  printf("{\"test\": \"%s\", \"ms\": %.3f, \"sm_mhz_effective\": %.0f, \"fp32_lane_ops_per_clk_per_sm\": %.2f, \"mufu_per_clk_per_sm\": %.2f, "
         "\"fp32_lane_ops_per_s\": %.4g, \"mufu_per_s\": %.4g}\n",
         name, ms, mhz, lane_ops_per_iter * iters * threads / cycles / sms, mufu_per_iter * iters * threads / cycles / sms,
         lane_ops_per_iter * iters * threads / (ms * 1e-3), mufu_per_iter * iters * threads / (ms * 1e-3));
  cudaFree(out); cudaFree(clk);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
  run<0>("ffma_3reg", 16, 0, sms);
  run<5>("ffma_imm", 16, 0, sms);
  run<1>("ffma2", 16, 0, sms);
  run<2>("mufu_ex2", 0, 8, sms);
  run<3>("mix_13ffma_1mufu", 13, 1, sms);
  run<4>("mix_13ffma2_2mufu", 28, 2, sms);
  return 0;
}
