// mix.cu -- does packing FP32 work as f32x2 (FFMA2) buy ISSUE slots on B200 when the FP32 pipe is not the limiter?
// The fused row kernels are instruction-issue bound with the FMA pipe at 30-40 %: ~45 % of their
// instructions are FFMA/FMUL/FADD, the rest integer / select / conversion work.  Each thread here carries TWO
// independent "rows" of the same instruction mix per iteration.  Template parameter NF = FMAs per row and iteration
// (12: FP32-heavy, 63 % of the slots; 6: the row kernels' share, ~45 %), next to 8 integer/select ops + 1 MUFU per row:
//   mode 0: scalar,   mode 1: the FMAs packed as FFMA2,   mode 2 / 3: the FMAs alone, scalar / packed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu ; run: ./mix
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE, int NF>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float2 a[6], b[6];
  unsigned h0 = threadIdx.x * 2654435761u, h1 = h0 ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < 6; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, 0.5f + i); b[i] = make_float2(1.0f + i * 1e-3f, 1.0f - i * 1e-3f); }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { a[i].x = fmaf(a[i].x, b[i].x, b[(i + 1) % 6].x); a[i].y = fmaf(a[i].y, b[i].y, b[(i + 1) % 6].y); }
      if (NF > 6) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { b[i].x = fmaf(b[i].x, a[i].x, a[(i + 3) % 6].x); b[i].y = fmaf(b[i].y, a[i].y, a[(i + 3) % 6].y); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = __ffma2_rn(a[i], b[i], b[(i + 1) % 6]);
      if (NF > 6) {
#pragma unroll
        for (int i = 0; i < 6; ++i) b[i] = __ffma2_rn(b[i], a[i], a[(i + 3) % 6]);
      }
    }
    if (MODE <= 1) {
      // 6 integer / select ops per row (Philox-like mixing + a data-dependent select) and one MUFU per row
      h0 = (h0 ^ (h0 >> 13)) + 0x85ebca6bu; h1 = (h1 ^ (h1 >> 13)) + 0x85ebca6bu;
      h0 ^= h0 << 7; h1 ^= h1 << 7;
      h0 = (h0 & 0x7fffffffu) | (h1 >> 31); h1 = (h1 & 0x7fffffffu) | (h0 >> 31);
      h0 += 0x9e3779b9u; h1 += 0x9e3779b9u;
      a[0].x = (h0 & 0x10000u) ? a[0].x : b[0].x; a[0].y = (h1 & 0x10000u) ? a[0].y : b[0].y;
      float r0, r1;
      asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(fabsf(a[1].x) + 1.0f));
      asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(fabsf(a[1].y) + 1.0f));
      b[5].x = r0; b[5].y = r1;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += a[i].x + a[i].y + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(h0 ^ h1);
}

template <int MODE, int NF>
float run(float* out, int grid, int iters) {
  k<MODE, NF><<<grid, 256>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, NF><<<grid, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 4 * 256);
  const int grid = sms * 4, iters = 20000;   // 4 CTAs x 8 warps per SM, like the row kernels
  {
    const float t0 = run<0, 12>(out, grid, iters), t1 = run<1, 12>(out, grid, iters), t2 = run<2, 12>(out, grid, iters), t3 = run<3, 12>(out, grid, iters);
    printf("{\"fma_per_row\": 12, \"mixed_scalar_ms\": %.3f, \"mixed_packed_ms\": %.3f, \"packed_over_scalar\": %.3f, \"fma_only_scalar_ms\": %.3f, \"fma_only_packed_ms\": %.3f}\n",
           t0, t1, t1 / t0, t2, t3);
  }
  {
    const float t0 = run<0, 6>(out, grid, iters), t1 = run<1, 6>(out, grid, iters), t2 = run<2, 6>(out, grid, iters), t3 = run<3, 6>(out, grid, iters);
    printf("{\"fma_per_row\": 6, \"mixed_scalar_ms\": %.3f, \"mixed_packed_ms\": %.3f, \"packed_over_scalar\": %.3f, \"fma_only_scalar_ms\": %.3f, \"fma_only_packed_ms\": %.3f}\n",
           t0, t1, t1 / t0, t2, t3);
  }
  return 0;
}
