// greenctx_probe.cu — can this driver split the B200's SMs into two green contexts, do runtime-API
// launches on their streams stay inside their partition, and do the two partitions run concurrently?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o greenctx_probe greenctx_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <set>
#include <vector>

#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); \
  printf("%s -> %d %s\n", #x, (int)r_, s_ ? s_ : "?"); return 1; } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)

__global__ void k_smid(int *out, long long spin) {
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
}

int main() {
  RT(cudaSetDevice(0));
  RT(cudaFree(0));
  CU(cuInit(0));
  CUdevice dev;
  CU(cuDeviceGet(&dev, 0));
  CUdevResource all;
  CU(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  unsigned int ng = 0;
  CU(cuDevSmResourceSplitByCount(nullptr, &ng, &all, nullptr, 0, 8));
  printf("groups of >= 8 SMs: %u\n", ng);
  std::vector<CUdevResource> groups(ng);
  CUdevResource rem;
  CU(cuDevSmResourceSplitByCount(groups.data(), &ng, &all, &rem, 0, 8));
  printf("remainder SMs: %u, group sizes:", rem.sm.smCount);
  for (unsigned i = 0; i < ng; i++) printf(" %u", groups[i].sm.smCount);
  printf("\n");
  const unsigned laneGroups = 6;   // 48 SMs
  std::vector<CUdevResource> ra(groups.begin(), groups.begin() + laneGroups), rb(groups.begin() + laneGroups, groups.end());
  if (rem.sm.smCount > 0) rb.push_back(rem);
  CUdevResourceDesc da, db;
  CU(cuDevResourceGenerateDesc(&da, ra.data(), (unsigned)ra.size()));
  CU(cuDevResourceGenerateDesc(&db, rb.data(), (unsigned)rb.size()));
  CUgreenCtx ga, gb;
  CU(cuGreenCtxCreate(&ga, da, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(cuGreenCtxCreate(&gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sa, sb;
  CU(cuGreenCtxStreamCreate(&sa, ga, CU_STREAM_NON_BLOCKING, 0));
  CU(cuGreenCtxStreamCreate(&sb, gb, CU_STREAM_NON_BLOCKING, -1));
  const int nb = 2000;
  int *oa, *ob;
  RT(cudaMalloc(&oa, nb * sizeof(int)));
  RT(cudaMalloc(&ob, nb * sizeof(int)));
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
  const long long spin = 200000;   // ~0.1 ms per block
  // warm-up + membership
  k_smid<<<nb, 64, 0, (cudaStream_t)sa>>>(oa, 1000);
  k_smid<<<nb, 64, 0, (cudaStream_t)sb>>>(ob, 1000);
  RT(cudaDeviceSynchronize());
  std::vector<int> ha(nb), hb(nb);
  RT(cudaMemcpy(ha.data(), oa, nb * sizeof(int), cudaMemcpyDeviceToHost));
  RT(cudaMemcpy(hb.data(), ob, nb * sizeof(int), cudaMemcpyDeviceToHost));
  std::set<int> sa_set(ha.begin(), ha.end()), sb_set(hb.begin(), hb.end());
  int common = 0;
  for (int x : sa_set) common += sb_set.count(x);
  printf("partition A used %zu SMs, partition B used %zu SMs, in common %d\n", sa_set.size(), sb_set.size(), common);
  // concurrency: each alone, then both
  float ta, tb, tab;
  cudaEventRecord(e0, (cudaStream_t)sa); k_smid<<<480, 64, 0, (cudaStream_t)sa>>>(oa, spin); cudaEventRecord(e1, (cudaStream_t)sa);
  RT(cudaDeviceSynchronize()); cudaEventElapsedTime(&ta, e0, e1);
  cudaEventRecord(e2, (cudaStream_t)sb); k_smid<<<1000, 64, 0, (cudaStream_t)sb>>>(ob, spin); cudaEventRecord(e3, (cudaStream_t)sb);
  RT(cudaDeviceSynchronize()); cudaEventElapsedTime(&tb, e2, e3);
  cudaEventRecord(e0, (cudaStream_t)sa); cudaEventRecord(e2, (cudaStream_t)sb);
  k_smid<<<480, 64, 0, (cudaStream_t)sa>>>(oa, spin);
  k_smid<<<1000, 64, 0, (cudaStream_t)sb>>>(ob, spin);
  cudaEventRecord(e1, (cudaStream_t)sa); cudaEventRecord(e3, (cudaStream_t)sb);
  RT(cudaDeviceSynchronize());
  float t1, t2;
  cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e2, e3);
  printf("alone: A %.3f ms, B %.3f ms; together: A %.3f ms, B %.3f ms\n", ta, tb, t1, t2);
  // cross-stream event dependency between the partitions
  cudaEventRecord(e0, (cudaStream_t)sa);
  RT(cudaStreamWaitEvent((cudaStream_t)sb, e0, 0));
  k_smid<<<10, 64, 0, (cudaStream_t)sb>>>(ob, 1000);
  RT(cudaDeviceSynchronize());
  printf("ok\n");
  return 0;
}
