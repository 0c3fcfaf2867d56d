// Microbenchmark: dependent-issue latency (cycles per op, one warp, one chain) of the operations on
// the serial chains of the lane kernels (pilot PLL, RDS Costas loop, AGC) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../fmtuner_sdr_b200/csrc \
//        -o lat_bench lat_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fm_math.h"

__device__ __forceinline__ uint32_t ncoConstrainDev(float theta) {
  const float p = (float)((double)theta * 0.159154943091895);
  float fpart = p - truncf(p);
  if (fpart < 0.0f) fpart = fpart + 1.0f;
  const float scaled = fpart * 4294967296.0f;
  return (scaled >= 4294967296.0f) ? 0u : __float2uint_rz(scaled);
}
__device__ __forceinline__ float ncoPhaseDev(uint32_t theta) {
  return (float)(6.283185307179586 * (double)((float)theta) / 4294967296.0);
}

constexpr int N = 4096;

#define CHAIN(NAME, DECL, BODY, SINK)                                              \
  __global__ void k_##NAME(float *out, long long *cyc, float a, float b, int n) {   \
    DECL;                                                                          \
    long long t0 = clock64();                                                      \
    for (int i = 0; i < n; i += 8) {                                               \
      BODY BODY BODY BODY BODY BODY BODY BODY                                      \
    }                                                                              \
    long long t1 = clock64();                                                      \
    out[threadIdx.x] = SINK;                                                       \
    if (threadIdx.x == 0) *cyc = t1 - t0;                                          \
  }

CHAIN(fmul, float x = a + threadIdx.x, x = __fmul_rn(x, b);, x)
CHAIN(ffma, float x = a + threadIdx.x, x = __fmaf_rn(x, b, a);, x)
CHAIN(fadd, float x = a + threadIdx.x, x = __fadd_rn(x, b);, x)
CHAIN(f2f_roundtrip, float x = a + threadIdx.x, x = (float)((double)x);asm volatile("" : "+f"(x));, x)
CHAIN(dmul_rt, float x = a + threadIdx.x, x = (float)((double)x * 1.0000001);, x)
CHAIN(dmul, double x = a + threadIdx.x, x = x * (double)b;, (float)x)
CHAIN(dfma, double x = a + threadIdx.x, x = fma(x, (double)b, (double)a);, (float)x)
CHAIN(trunc, float x = a + threadIdx.x, x = truncf(x) + b;, x)
CHAIN(rint, float x = a + threadIdx.x, x = rintf(x) + b;, x)
CHAIN(f2i_i2f, float x = a + threadIdx.x, x = (float)__float2uint_rz(x) + b;, x)
CHAIN(f2i_rn_i2f, float x = a + threadIdx.x, x = (float)__float2int_rn(x) + b;, x)
CHAIN(imad, uint32_t x = (uint32_t)a + threadIdx.x; uint32_t bb = (uint32_t)b, x = x * bb + 3u;, (float)x)
CHAIN(imad_wide, unsigned long long x = (unsigned long long)a + threadIdx.x; uint32_t bb = (uint32_t)b,
      x = (unsigned long long)(uint32_t)x * bb + (x >> 32);, (float)x)
CHAIN(mulhi, uint32_t x = (uint32_t)a + threadIdx.x; uint32_t bb = (uint32_t)b + 0x9e3779b9u, x = __umulhi(x, bb) + bb;, (float)x)
CHAIN(sel, float x = a + threadIdx.x, x = (x > b) ? x - b : x + a;, x)
CHAIN(fdiv, float x = a + threadIdx.x, x = __fdiv_rn(x, b) + a;, x)
CHAIN(fsqrt, float x = a + threadIdx.x, x = __fsqrt_rn(x) + a;, x)
CHAIN(rcp_approx, float x = a + threadIdx.x, x = __frcp_rn(x) + a;, x)
CHAIN(constrain, float x = a * 1e-3f + threadIdx.x * 1e-6f, x = __uint_as_float((ncoConstrainDev(x) >> 9) | 0x3a000000u);, x)
CHAIN(phase, uint32_t x = (uint32_t)(a * 1e6f) + threadIdx.x, x = __float_as_uint(ncoPhaseDev(x)) * 77u;, (float)x)
CHAIN(sincos, float x = a + threadIdx.x * 0.01f, { float s_; float c_; fm_sincosf(x, &s_, &c_); x = s_ + c_; }, x)
CHAIN(pllstep, uint32_t th = (uint32_t)(a * 1e6f) + threadIdx.x * 1000u; uint32_t dth = 340000000u; float vq = 0.1f; float ph = 0.f,
      { const float err = b * vq; dth += ncoConstrainDev(err * 0.01f); th += ncoConstrainDev(err * 0.1f); th += dth;
        ph = ncoPhaseDev(th); float s_; float c_; fm_sincosf(ph, &s_, &c_); vq = s_; }, vq + ph)

__global__ void k_lds(float *out, long long *cyc, float a, float b, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i * 33 + 32) & 1023;
  __syncwarp();
  int x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = idx[x];
  long long t1 = clock64();
  out[threadIdx.x] = (float)x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

#define RUN(NAME)                                                        \
  k_##NAME<<<1, 32>>>(out, cyc, 1.25f, 1.0000001f, N);                   \
  k_##NAME<<<1, 32>>>(out, cyc, 1.25f, 1.0000001f, N);                   \
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                \
  printf("%-16s %8.2f cycles/iter\n", #NAME, (double)h / N);

int main() {
  float *out;
  long long *cyc, h = 0;
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, 8);
  RUN(fmul) RUN(ffma) RUN(fadd) RUN(f2f_roundtrip) RUN(dmul_rt) RUN(dmul) RUN(dfma) RUN(trunc) RUN(rint)
  RUN(f2i_i2f) RUN(f2i_rn_i2f) RUN(imad) RUN(imad_wide) RUN(mulhi) RUN(sel) RUN(fdiv) RUN(fsqrt) RUN(rcp_approx)
  RUN(constrain) RUN(phase) RUN(sincos) RUN(pllstep) RUN(lds)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
