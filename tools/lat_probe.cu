// lat_probe.cu -- development probe: latencies seen by a 16-CTA cluster (the select kernel's shape) in its
// green-context SM partition, idle and while a streaming kernel saturates HBM from the other partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/lat_probe tools/lat_probe.cu
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("FAIL %s: %d\n", #x, (int)r_); return 1; } } while (0)

constexpr int NT = 256, NCTA = 16, ITER = 64;

__global__ void k_stream(const double2* __restrict__ a, double2* __restrict__ b, size_t n, int reps) {
  for (int r = 0; r < reps; r++)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      double2 v = a[i];
      v.x += 1.0;
      b[i] = v;
    }
}

__device__ __forceinline__ double chain32(double x, const double* f, const double* p) {
#pragma unroll
  for (int u = 0; u < 32; u++) x = __dsub_rn(x, __dmul_rn(f[u], p[u]));
  return x;
}

// res[k] = average clocks of pattern k (thread 0 of CTA 0)
__global__ void __launch_bounds__(NT, 1) k_probe(const double* __restrict__ big, size_t ld, int rows, double* xbuf,
                                                  long long* res, double* sink) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ double sm[];
  const int tid = threadIdx.x, crank = (int)cl.block_rank();
  const int gid = crank * NT + tid;
  const bool rec = (gid == 0);
  double acc = 0.0;
  long long t0, t1;
  sm[tid] = tid;
  __syncthreads();
  // 0: cluster.sync alone
  cl.sync();
  t0 = clock64();
  for (int it = 0; it < ITER; it++) cl.sync();
  t1 = clock64();
  if (rec) res[0] = (t1 - t0) / ITER;
  // 1: one coalesced global store per thread, then cluster.sync
  t0 = clock64();
  for (int it = 0; it < ITER; it++) {
    xbuf[gid] = (double)it;
    cl.sync();
  }
  t1 = clock64();
  if (rec) res[1] = (t1 - t0) / ITER;
  // 2: load (ld.cg) of a value a peer CTA stored before the barrier: issue -> use
  long long sum = 0;
  for (int it = 0; it < ITER; it++) {
    xbuf[4096 + gid] = (double)(it + gid);
    cl.sync();
    const int peer = ((crank + 1) % NCTA) * NT + tid;
    t0 = clock64();
    const double v = __ldcg(xbuf + 4096 + peer);
    acc += v;
    if (acc == -1.0) res[63] = 1;  // dependency
    t1 = clock64();
    sum += t1 - t0;
    cl.sync();
  }
  if (rec) res[2] = sum / ITER;
  // 3: cold DRAM gather, one 8-byte element per thread, row stride ld (a fresh column each iteration)
  sum = 0;
  for (int it = 0; it < ITER; it++) {
    const size_t col = 1000 + 37 * it;
    t0 = clock64();
    const double v = big[(size_t)(gid % rows) * ld + col];
    acc += v;
    if (acc == -1.0) res[63] = 1;
    t1 = clock64();
    sum += t1 - t0;
    __syncthreads();
  }
  if (rec) res[3] = sum / ITER;
  // 4: cold DRAM contiguous row read, 3 elements per thread (coalesced)
  sum = 0;
  for (int it = 0; it < ITER; it++) {
    const size_t row = (size_t)(17 * it + 5) % rows;
    t0 = clock64();
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) v += big[row * ld + 6000 + gid + c * 4096 - 6000 * (c > 0)];
    acc += v;
    if (acc == -1.0) res[63] = 1;
    t1 = clock64();
    sum += t1 - t0;
    __syncthreads();
  }
  if (rec) res[4] = sum / ITER;
  // 5: DSMEM load from a peer CTA
  sum = 0;
  cl.sync();
  for (int it = 0; it < ITER; it++) {
    const double* rp = cl.map_shared_rank(sm, (crank + 1 + it) % NCTA);
    t0 = clock64();
    const double v = rp[(tid + it) % NT];
    acc += v;
    if (acc == -1.0) res[63] = 1;
    t1 = clock64();
    sum += t1 - t0;
  }
  cl.sync();
  if (rec) res[5] = sum / ITER;
  // 6: __syncthreads
  t0 = clock64();
  for (int it = 0; it < ITER; it++) __syncthreads();
  t1 = clock64();
  if (rec) res[6] = (t1 - t0) / ITER;
  // 7: chain of 32 dependent (DMUL, DADD) pairs, operands in registers
  {
    double f[32], p[32];
#pragma unroll
    for (int u = 0; u < 32; u++) { f[u] = 1.0 + 1e-3 * (u + tid); p[u] = 1e-3 * (u + 1); }
    double x = (double)tid;
    t0 = clock64();
    for (int it = 0; it < ITER; it++) x = chain32(x, f, p);
    t1 = clock64();
    acc += x;
    if (rec) res[7] = (t1 - t0) / ITER;
    // 8: three independent chains (ILP 3)
    double x1 = x + 1, x2 = x + 2, x3 = x + 3;
    t0 = clock64();
    for (int it = 0; it < ITER; it++) { x1 = chain32(x1, f, p); x2 = chain32(x2, f, p); x3 = chain32(x3, f, p); }
    t1 = clock64();
    acc += x1 + x2 + x3;
    if (rec) res[8] = (t1 - t0) / ITER;
    // 9: DDIV
    double d = 3.0 + tid;
    t0 = clock64();
    for (int it = 0; it < ITER; it++) d = __ddiv_rn(d + 1.0, 1.0000001 + x);
    t1 = clock64();
    acc += d;
    if (rec) res[9] = (t1 - t0) / ITER;
  }
  // 10: warp argmin with 3 redux + ballot
  {
    unsigned k1 = tid * 2654435761u, k2 = gid * 40503u;
    int idx = gid;
    t0 = clock64();
    for (int it = 0; it < ITER; it++) {
      const unsigned m1 = __reduce_min_sync(0xffffffffu, k1);
      const unsigned m2 = __reduce_min_sync(0xffffffffu, (k1 == m1) ? k2 : 0xffffffffu);
      const unsigned m3 = __reduce_min_sync(0xffffffffu, (k1 == m1 && k2 == m2) ? (unsigned)idx : 0xffffffffu);
      const unsigned own = __ballot_sync(0xffffffffu, (unsigned)idx == m3);
      k1 = k1 * 1664525u + m1 + own;
      k2 += m2 + m3;
    }
    t1 = clock64();
    acc += k1 + k2;
    if (rec) res[10] = (t1 - t0) / ITER;
  }
  // 11: DSMEM store to every peer (lane r -> CTA r) + cluster.sync + local read of 128 entries by a warp
  {
    double* slots = sm + 1024;
    t0 = clock64();
    for (int it = 0; it < ITER; it++) {
      const int lane = tid & 31, w = tid >> 5;
      if (lane < NCTA) cl.map_shared_rank(slots, lane)[(crank * 8 + w) * 4] = (double)(it + gid);
      cl.sync();
      double m = 1e300;
      for (int k = lane; k < 128; k += 32) m = fmin(m, slots[k * 4]);
      acc += m;
      cl.sync();
    }
    t1 = clock64();
    if (rec) res[11] = (t1 - t0) / ITER;
  }
  // 12: 4 coalesced global stores per thread, NO barrier: time to issue only
  t0 = clock64();
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int c = 0; c < 4; c++) xbuf[8192 + c * 4096 + gid] = (double)it;
  t1 = clock64();
  if (rec) res[12] = (t1 - t0) / ITER;
  // 13: L2-resident load (data written long ago by this same kernel, other CTA), no barrier in between
  cl.sync();
  sum = 0;
  for (int it = 0; it < ITER; it++) {
    const int peer = ((crank + 3) % NCTA) * NT + ((tid + it) % NT);
    t0 = clock64();
    const double v = __ldcg(xbuf + 8192 + peer);
    acc += v;
    if (acc == -1.0) res[63] = 1;
    t1 = clock64();
    sum += t1 - t0;
  }
  if (rec) res[13] = sum / ITER;
  cl.sync();
  if (acc == 12345.678) *sink = acc;
}

template <class F>
static bool entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  *fn = (F)p;
  return true;
}

int main() {
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
  CUdevResource in, grp[1], rem;
  unsigned n = 1;
  CU(pGetRes(dev, &in, CU_DEV_RESOURCE_TYPE_SM));
  CU(pSplit(grp, &n, &in, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE, 16));
  CUdevResourceDesc dA, dB;
  CU(pDesc(&dA, &grp[0], 1));
  CU(pDesc(&dB, &rem, 1));
  CUgreenCtx gA, gB;
  CU(pCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(pCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sA, sB;
  CU(pStream(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
  CU(pStream(&sB, gB, CU_STREAM_NON_BLOCKING, 0));

  const size_t ld = 12304, rows = 4097;
  double *big, *big2, *xbuf, *sink;
  long long* res;
  CK(cudaMalloc(&big, ld * rows * 8));
  CK(cudaMalloc(&big2, ld * rows * 8));
  CK(cudaMemset(big, 0, ld * rows * 8));
  CK(cudaMalloc(&xbuf, 65536 * 8));
  CK(cudaMemset(xbuf, 0, 65536 * 8));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMalloc(&res, 64 * 8));
  const size_t smem = 190 * 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const char* names[14] = {"cluster.sync", "global store + cluster.sync", "ld.cg of peer-written line (after barrier)",
                           "cold DRAM gather (1 elt/thread, row stride)", "cold DRAM row read (3 elts/thread)",
                           "DSMEM load from peer", "__syncthreads", "chain of 32 (DMUL,DADD)", "3 chains of 32 (ILP 3)",
                           "DDIV (dependent)", "warp argmin (3 redux + ballot)",
                           "DSMEM publish + cluster.sync + collect + cluster.sync", "issue 4 global stores/thread",
                           "ld.cg of L2-resident line (no barrier)"};
  for (int load = 0; load < 2; load++) {
    CK(cudaMemset(res, 0, 64 * 8));
    if (load) {
      k_stream<<<132 * 8, 256, 0, (cudaStream_t)sB>>>((const double2*)big2, (double2*)big2, ld * rows / 2 / 2, 40);
      CK(cudaGetLastError());
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NCTA);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)sA;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NCTA;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, (cudaStream_t)sA));
    CK(cudaLaunchKernelEx(&cfg, k_probe, (const double*)big, ld, (int)rows, xbuf, res, sink));
    CK(cudaEventRecord(e1, (cudaStream_t)sA));
    CK(cudaStreamSynchronize((cudaStream_t)sA));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaError_t se = cudaStreamQuery((cudaStream_t)sB);
    long long h[64];
    CK(cudaMemcpy(h, res, sizeof h, cudaMemcpyDeviceToHost));
    printf("=== %s (probe kernel %.1f us; stream kernel %s when the probe ended)\n",
           load ? "HBM saturated by the other partition" : "idle GPU", ms * 1e3,
           load ? (se == cudaErrorNotReady ? "still running" : "ALREADY DONE") : "-");
    for (int k = 0; k < 14; k++) printf("  %-58s %6lld clk\n", names[k], h[k]);
    CK(cudaDeviceSynchronize());
  }
  return 0;
}
