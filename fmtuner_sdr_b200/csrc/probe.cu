// probe.cu — measurement aids behind the C ABI (not on the data path).
//
// fmgpu_measure_fp32_tflops: the FP32 FMA roof the FIR kernels are held against in bench.py,
// MEASURED on the device at hand instead of derived from 148 SM x 128 lanes x 2 x clock: a
// register-resident packed-FMA (fma.rn.f32x2, the instruction the FIR kernels issue) loop on every
// SM, 8 independent accumulator pairs per thread, timed with CUDA events.
#include <cuda_runtime.h>

#include "../../include/fmgpu.h"

namespace {

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

__global__ void __launch_bounds__(256) k_fp32_peak(float2 *out, int iters, float h) {
  float2 acc[8], x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    acc[j] = make_float2(threadIdx.x * 1e-3f, j * 1e-3f);
    x[j] = out[(threadIdx.x + j) & 255];  // run-time values: nothing folds
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        acc[j] = ffma2(h, x[(j + u) & 7], acc[j]);
      }
    }
  }
  float2 s = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    s.x += acc[j].x;
    s.y += acc[j].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int fmgpu_measure_fp32_tflops(int device, double *tflops_out) {
  if (!tflops_out || cudaSetDevice(device) != cudaSuccess) {
    return FMGPU_ENODEV;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, iters = 4000;
  float2 *out = nullptr;
  if (cudaMalloc(&out, sizeof(float2) * blocks * 256) != cudaSuccess) {
    return FMGPU_ENOMEM;
  }
  cudaMemset(out, 0x3c, sizeof(float2) * blocks * 256);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {   // first repetition warms up
    cudaEventRecord(a);
    k_fp32_peak<<<blocks, 256>>>(out, iters, 0.999f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    const double fma = static_cast<double>(blocks) * 256.0 * iters * 128.0;  // 64 FFMA2 = 128 FMA per iteration
    if (rep > 0 && ms > 0.0f) {
      best = fmax(best, 2.0 * fma / (ms * 1e-3) / 1e12);
    }
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  *tflops_out = best;
  return cudaGetLastError() == cudaSuccess ? FMGPU_OK : FMGPU_ENODEV;
}
