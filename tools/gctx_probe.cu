// gctx_probe.cu -- development probe: can a 16-CTA cluster kernel run inside a small green-context SM
// partition while a big grid runs in the complementary partition?  (build: nvcc -arch=sm_100a -o gctx_probe)
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_));                         \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define CU(x)                                       \
  do {                                              \
    CUresult r_ = (x);                              \
    if (r_ != CUDA_SUCCESS) {                       \
      printf("FAIL %s: CUresult %d\n", #x, (int)r_); \
      return 1;                                     \
    }                                               \
  } while (0)

__global__ void k_cluster(int* smid, long long* t, int spin_us) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ double sm[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) {
    smid[blockIdx.x] = (int)id;
    t[2 * blockIdx.x] = t0;
  }
  sm[threadIdx.x] = 1.0;
  for (int it = 0; it < 50; it++) cl.sync();
  long long t1 = t0;
  while (t1 - t0 < (long long)spin_us * 1000) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  cl.sync();
  if (threadIdx.x == 0) t[2 * blockIdx.x + 1] = t1;
}

__global__ void k_big(int* smid, long long* t, int spin_us) {
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long t1 = t0;
  while (t1 - t0 < (long long)spin_us * 1000) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) {
    smid[blockIdx.x] = (int)id;
    t[2 * blockIdx.x] = t0;
    t[2 * blockIdx.x + 1] = t1;
  }
}

template <class F>
static bool entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
    printf("no driver entry point %s\n", name);
    return false;
  }
  *fn = (F)p;
  return true;
}

int main(int argc, char** argv) {
  const int want = argc > 1 ? atoi(argv[1]) : 16;
  const unsigned flags = argc > 2 ? (unsigned)atoi(argv[2]) : 0;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  decltype(&cuDeviceGetDevResource) pGetRes;
  decltype(&cuDevSmResourceSplitByCount) pSplit;
  decltype(&cuDevResourceGenerateDesc) pDesc;
  decltype(&cuGreenCtxCreate) pCreate;
  decltype(&cuGreenCtxStreamCreate) pStream;
  decltype(&cuDeviceGet) pDevGet;
  if (!entry("cuDeviceGetDevResource", &pGetRes) || !entry("cuDevSmResourceSplitByCount", &pSplit) ||
      !entry("cuDevResourceGenerateDesc", &pDesc) || !entry("cuGreenCtxCreate", &pCreate) ||
      !entry("cuGreenCtxStreamCreate", &pStream) || !entry("cuDeviceGet", &pDevGet))
    return 1;
  CUdevice dev;
  CU(pDevGet(&dev, 0));
  CUdevResource in;
  CU(pGetRes(dev, &in, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", in.sm.smCount);
  CUdevResource grp[1], rem;
  unsigned n = 1;
  CU(pSplit(grp, &n, &in, &rem, flags, (unsigned)want));
  printf("split(min=%d, flags=%u): groups=%u group0=%u SMs, remaining=%u SMs\n", want, flags, n, grp[0].sm.smCount,
         rem.sm.smCount);
  CUdevResourceDesc dA, dB;
  CU(pDesc(&dA, &grp[0], 1));
  CU(pDesc(&dB, &rem, 1));
  CUgreenCtx gA, gB;
  CU(pCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(pCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sA, sB;
  CU(pStream(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
  CU(pStream(&sB, gB, CU_STREAM_NON_BLOCKING, 0));

  int *dsA, *dsB;
  long long *dtA, *dtB;
  const int NB = 4096;
  CK(cudaMalloc(&dsA, 64 * 4));
  CK(cudaMalloc(&dsB, NB * 4));
  CK(cudaMalloc(&dtA, 64 * 16));
  CK(cudaMalloc(&dtB, NB * 16));
  for (int csize : {16, 8}) {
    for (size_t smem : {(size_t)200 * 1024, (size_t)100 * 1024}) {
      cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(csize);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)sA;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nact = -1;
      cudaError_t oe = cudaOccupancyMaxActiveClusters(&nact, k_cluster, &cfg);
      printf("cluster %2d smem %3zu KB in green ctx A: maxActiveClusters=%d (%s)\n", csize, smem >> 10, nact,
             cudaGetErrorString(oe));
      cfg.stream = 0;
      oe = cudaOccupancyMaxActiveClusters(&nact, k_cluster, &cfg);
      printf("   same on the default stream: maxActiveClusters=%d (%s)\n", nact, cudaGetErrorString(oe));
      cfg.stream = (cudaStream_t)sA;
      // big grid first (it would fill every SM it may use), then the cluster kernel
      CK(cudaMemset(dsA, 0xff, 64 * 4));
      k_big<<<NB, 256, 0, (cudaStream_t)sB>>>(dsB, dtB, 20);
      CK(cudaGetLastError());
      cudaError_t le = cudaLaunchKernelEx(&cfg, k_cluster, dsA, dtA, 100);
      printf("   launch: %s\n", cudaGetErrorString(le));
      cudaError_t se = cudaDeviceSynchronize();
      printf("   sync: %s\n", cudaGetErrorString(se));
      if (le != cudaSuccess || se != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      std::vector<int> hA(64), hB(NB);
      std::vector<long long> tA(128), tB(2 * NB);
      CK(cudaMemcpy(hA.data(), dsA, 64 * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hB.data(), dsB, NB * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(tA.data(), dtA, 64 * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(tB.data(), dtB, NB * 16, cudaMemcpyDeviceToHost));
      std::set<int> sa(hA.begin(), hA.begin() + csize), sb(hB.begin(), hB.end());
      long long b0 = tB[0], b1 = tB[1];
      for (int k = 0; k < NB; k++) {
        b0 = std::min(b0, tB[2 * k]);
        b1 = std::max(b1, tB[2 * k + 1]);
      }
      printf("   cluster SMs:");
      for (int x : sa) printf(" %d", x);
      int common = 0;
      for (int x : sa) common += sb.count(x);
      printf("\n   big grid used %zu SMs, %d shared with the cluster; cluster start %+lld us after big start, big ran %lld us\n",
             sb.size(), common, (tA[0] - b0) / 1000, (b1 - b0) / 1000);
    }
  }
  return 0;
}
