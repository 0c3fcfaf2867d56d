// Microbenchmark: FP32 FMA rate with scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float h, float2 x, float2 acc) {
  unsigned long long hh, xx, aa, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
  asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x.x), "f"(x.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(hh), "l"(xx), "l"(aa));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}

template <int PACKED>
__global__ void __launch_bounds__(256) k(float2 *out, int iters, float h0) {
  float2 acc[8];
  float2 x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    acc[j] = make_float2(threadIdx.x * 1e-3f, j * 1e-3f);
    x[j] = out[(threadIdx.x + j) & 255];  // runtime values: nothing folds
  }
  float h = h0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (PACKED) {
          acc[j] = ffma2(h, x[(j + u) & 7], acc[j]);
        } else {
          acc[j].x = fmaf(h, x[(j + u) & 7].x, acc[j].x);
          acc[j].y = fmaf(h, x[(j + u) & 7].y, acc[j].y);
        }
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    s.x += acc[j].x;
    s.y += acc[j].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 20000;
  float2 *out;
  cudaMalloc(&out, sizeof(float2) * blocks * 256);
  cudaMemset(out, 0x3c, sizeof(float2) * blocks * 256);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int packed = 0; packed < 2; packed++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(a);
      if (packed) {
        k<1><<<blocks, 256>>>(out, iters, 0.999f);
      } else {
        k<0><<<blocks, 256>>>(out, iters, 0.999f);
      }
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      const double fma = (double)blocks * 256 * iters * 128.0;
      printf("%s rep %d: %.3f ms, %.2f TFLOP/s (2 flop per FMA)\n", packed ? "FFMA2" : "FFMA ", rep, ms,
             2.0 * fma / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
