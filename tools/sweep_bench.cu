// sweep_bench.cu -- development microbenchmark: variants of the in-place rank-1 sweep
// T[i,j] -= f[i]*p[j] on the cfg2 shape, to pick the access pattern that gets closest to the
// measured copy bandwidth.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

// ---- plain copy / scale baselines ----
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n; i += 4 * st) {
    double2 x0 = a[i], x1 = a[i + st], x2 = a[i + 2 * st], x3 = a[i + 3 * st];
    b[i] = x0; b[i + st] = x1; b[i + 2 * st] = x2; b[i + 3 * st] = x3;
  }
  for (; i < n; i += st) b[i] = a[i];
}
__global__ void k_scale(double2* __restrict__ a, size_t n, double s) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n; i += 4 * st) {
    double2 x0 = a[i], x1 = a[i + st], x2 = a[i + 2 * st], x3 = a[i + 3 * st];
    x0.x *= s; x0.y *= s; x1.x *= s; x1.y *= s; x2.x *= s; x2.y *= s; x3.x *= s; x3.y *= s;
    a[i] = x0; a[i + st] = x1; a[i + 2 * st] = x2; a[i + 3 * st] = x3;
  }
  for (; i < n; i += st) { double2 x = a[i]; x.x *= s; x.y *= s; a[i] = x; }
}

// ---- V0: column-fixed threads, column-group-major units (current product kernel) ----
template <int U>
__global__ void __launch_bounds__(256) k_v0(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p) {
  const int ldv = ld >> 1, nfull = ldv / 256;
  const long long Un = (long long)nfull * R;
  long long u0 = Un * blockIdx.x / gridDim.x; const long long u1 = Un * (blockIdx.x + 1) / gridDim.x;
  double2* T2 = (double2*)T; const double2* p2 = (const double2*)p;
  while (u0 < u1) {
    int cg = (int)(u0 / R), r0 = (int)(u0 - (long long)cg * R), r1 = (int)min((long long)R, r0 + (u1 - u0));
    int chunk = cg * 256 + threadIdx.x; double2 pr = p2[chunk]; double2* d = T2 + chunk;
    for (int r = r0; r < r1; r += U) {
      double2 x[U]; double fv[U];
#pragma unroll
      for (int k = 0; k < U; k++) if (r + k < r1) { fv[k] = f[r + k]; x[k] = d[(size_t)(r + k) * ldv]; }
#pragma unroll
      for (int k = 0; k < U; k++) if (r + k < r1) { double2 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv[k], pr.x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv[k], pr.y)); d[(size_t)(r + k) * ldv] = y; }
    }
    u0 += r1 - r0;
  }
}
// ---- V1: flat contiguous: the matrix is one long array of double2; p read through L1 (__ldg) ----
template <int U>
__global__ void __launch_bounds__(256) k_v1(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p) {
  const int ldv = ld >> 1; const size_t n = (size_t)R * ldv;
  double2* T2 = (double2*)T; const double2* p2 = (const double2*)p;
  // each CTA owns a contiguous span; within it threads stride by 256
  size_t per = (n + gridDim.x - 1) / gridDim.x; per = (per + 255) / 256 * 256;
  size_t s0 = per * blockIdx.x, s1 = min(n, s0 + per);
  for (size_t i = s0 + threadIdx.x; i < s1; i += 256 * U) {
    double2 x[U]; double fv[U]; double2 pr[U];
#pragma unroll
    for (int k = 0; k < U; k++) { size_t q = i + (size_t)k * 256; if (q < s1) { x[k] = T2[q]; int row = (int)(q / ldv), c = (int)(q - (size_t)row * ldv); fv[k] = __ldg(f + row); pr[k] = __ldg(p2 + c); } }
#pragma unroll
    for (int k = 0; k < U; k++) { size_t q = i + (size_t)k * 256; if (q < s1) { double2 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv[k], pr[k].x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv[k], pr[k].y)); T2[q] = y; } }
  }
}
// ---- V1m: V1 with optional full mirror (tile order reversed) to test L2 carry-over between sweeps ----
template <int U>
__global__ void __launch_bounds__(256) k_v1m(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p, int rev) {
  const unsigned ldv = ld >> 1; const unsigned long long n = (unsigned long long)R * ldv;
  double2* T2 = (double2*)T; const double2* p2 = (const double2*)p;
  const unsigned TILE = 256 * U; const unsigned long long nt = (n + TILE - 1) / TILE;
  unsigned long long t0 = nt * blockIdx.x / gridDim.x, t1 = nt * (blockIdx.x + 1) / gridDim.x;
  for (unsigned long long tt = t0; tt < t1; tt++) {
    unsigned long long t = rev ? nt - 1 - tt : tt; unsigned long long q0 = t * TILE + threadIdx.x;
    unsigned row = (unsigned)(q0 / ldv), c = (unsigned)(q0 - (unsigned long long)row * ldv);
    double2 x[U]; unsigned rw[U], cc[U];
#pragma unroll
    for (int k = 0; k < U; k++) { unsigned long long q = q0 + k * 256; rw[k] = row; cc[k] = c; if (q < n) x[k] = T2[q]; c += 256; while (c >= ldv) { c -= ldv; row++; } }
#pragma unroll
    for (int k = 0; k < U; k++) { unsigned long long q = q0 + k * 256; if (q < n) { double fv = __ldg(f + rw[k]); double2 pr = __ldg(p2 + cc[k]); double2 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv, pr.x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv, pr.y)); T2[q] = y; } }
  }
}
// ---- V1h: V1m with cache hints.  MODE 1: tableau loads bypass L1 (ld.global.L1::no_allocate) so prow/f stay
// L1 resident.  MODE 2: additionally L2 eviction policies: tiles written in the last `keep` fraction of the
// sweep get evict_last (they are the first ones the next, mirrored sweep reads), the others evict_first.
__device__ __forceinline__ double2 ld_na(const double2* p) {
  double2 r; asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r;
}
__device__ __forceinline__ double2 ld_pol(const double2* p, unsigned long long pol) {
  double2 r; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol)); return r;
}
__device__ __forceinline__ void st_pol(double2* p, double2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
template <int U, int MODE>
__global__ void __launch_bounds__(256) k_v1h(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p, int rev, float keep) {
  const unsigned ldv = ld >> 1; const unsigned long long n = (unsigned long long)R * ldv;
  double2* T2 = (double2*)T; const double2* p2 = (const double2*)p;
  const unsigned TILE = 256 * U; const unsigned long long nt = (n + TILE - 1) / TILE;
  unsigned long long t0 = nt * blockIdx.x / gridDim.x, t1 = nt * (blockIdx.x + 1) / gridDim.x;
  unsigned long long pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  for (unsigned long long tt = t0; tt < t1; tt++) {
    unsigned long long t = rev ? nt - 1 - tt : tt; unsigned long long q0 = t * TILE + threadIdx.x;
    // position of this tile in sweep order (0 = first processed, 1 = last)
    const bool late = (double)tt >= (1.0 - keep) * (double)nt;
    unsigned row = (unsigned)(q0 / ldv), c = (unsigned)(q0 - (unsigned long long)row * ldv);
    double2 x[U]; unsigned rw[U], cc[U];
#pragma unroll
    for (int k = 0; k < U; k++) { unsigned long long q = q0 + k * 256; rw[k] = row; cc[k] = c; if (q < n) x[k] = (MODE == 2) ? ld_pol(T2 + q, pol_first) : ld_na(T2 + q); c += 256; while (c >= ldv) { c -= ldv; row++; } }
#pragma unroll
    for (int k = 0; k < U; k++) { unsigned long long q = q0 + k * 256; if (q < n) { double fv = __ldg(f + rw[k]); double2 pr = __ldg(p2 + cc[k]); double2 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv, pr.x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv, pr.y));
      if (MODE == 2) st_pol(T2 + q, y, late ? pol_last : pol_first); else T2[q] = y; } }
  }
}
// ---- V2: row-major units: CTA takes (row, colgroup) units with colgroup fastest; p from L1 ----
template <int U>
__global__ void __launch_bounds__(256) k_v2(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p) {
  const int ldv = ld >> 1; const int ncg = (ldv + 255) / 256;
  const long long Un = (long long)ncg * R;
  long long u0 = Un * blockIdx.x / gridDim.x; const long long u1 = Un * (blockIdx.x + 1) / gridDim.x;
  double2* T2 = (double2*)T; const double2* p2 = (const double2*)p;
  for (long long u = u0; u < u1; u += U) {
    double2 x[U]; double fv[U]; double2 pr[U]; bool ok[U];
#pragma unroll
    for (int k = 0; k < U; k++) { long long q = u + k; ok[k] = false; if (q < u1) { int row = (int)(q / ncg), cg = (int)(q - (long long)row * ncg); int c = cg * 256 + threadIdx.x; if (c < ldv) { ok[k] = true; x[k] = T2[(size_t)row * ldv + c]; fv[k] = __ldg(f + row); pr[k] = __ldg(p2 + c); } } }
#pragma unroll
    for (int k = 0; k < U; k++) { long long q = u + k; if (ok[k]) { int row = (int)(q / ncg), cg = (int)(q - (long long)row * ncg); int c = cg * 256 + threadIdx.x; double2 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv[k], pr[k].x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv[k], pr[k].y)); T2[(size_t)row * ldv + c] = y; } }
  }
}
// ---- V3: column-fixed like V0 but 32 B per thread (two double2 -> 512 columns per 128 threads) ----
template <int U>
__global__ void __launch_bounds__(128) k_v3(double* T, int ld, int R, const double* __restrict__ f, const double* __restrict__ p) {
  const int ldq = ld >> 2, nfull = ldq / 128;  // double4 chunks
  const long long Un = (long long)nfull * R;
  long long u0 = Un * blockIdx.x / gridDim.x; const long long u1 = Un * (blockIdx.x + 1) / gridDim.x;
  double4* T4 = (double4*)T; const double4* p4 = (const double4*)p;
  while (u0 < u1) {
    int cg = (int)(u0 / R), r0 = (int)(u0 - (long long)cg * R), r1 = (int)min((long long)R, r0 + (u1 - u0));
    int chunk = cg * 128 + threadIdx.x; double4 pr = p4[chunk]; double4* d = T4 + chunk;
    for (int r = r0; r < r1; r += U) {
      double4 x[U]; double fv[U];
#pragma unroll
      for (int k = 0; k < U; k++) if (r + k < r1) { fv[k] = f[r + k]; x[k] = d[(size_t)(r + k) * ldq]; }
#pragma unroll
      for (int k = 0; k < U; k++) if (r + k < r1) { double4 y; y.x = __dsub_rn(x[k].x, __dmul_rn(fv[k], pr.x)); y.y = __dsub_rn(x[k].y, __dmul_rn(fv[k], pr.y)); y.z = __dsub_rn(x[k].z, __dmul_rn(fv[k], pr.z)); y.w = __dsub_rn(x[k].w, __dmul_rn(fv[k], pr.w)); d[(size_t)(r + k) * ldq] = y; }
    }
    u0 += r1 - r0;
  }
}

template <class F> float timeit(F launch, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char** argv) {
  int R = 4097, C = 12289; int ld = (C + 15) / 16 * 16;
  if (argc > 2) { R = atoi(argv[1]); C = atoi(argv[2]); ld = (C + 15) / 16 * 16; }
  int ld64 = (C + 63) / 64 * 64;
  size_t n = (size_t)R * ld64;
  double *T, *T2, *f, *p;
  CK(cudaMalloc(&T, n * 8)); CK(cudaMalloc(&T2, n * 8)); CK(cudaMalloc(&f, R * 8)); CK(cudaMalloc(&p, ld64 * 8));
  CK(cudaMemset(T, 0, n * 8)); CK(cudaMemset(T2, 0, n * 8));
  std::vector<double> hf(R, 1e-3), hp(ld64, 1e-3);
  CK(cudaMemcpy(f, hf.data(), R * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(p, hp.data(), ld64 * 8, cudaMemcpyHostToDevice));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double bytes = 16.0 * R * C;  // algorithmic
  const int reps = 20;
  auto rep = [&](const char* name, float ms) { printf("%-34s %8.2f us  %7.0f GB/s (algorithmic 16RC)\n", name, ms * 1e3, bytes / ms / 1e6); fflush(stdout); };
  size_t nv = (size_t)R * ld / 2;
  for (int g : {sms * 4, sms * 8, sms * 16, sms * 32}) { char nm[64];
    snprintf(nm, 64, "copy g=%d", g); rep(nm, timeit([&] { k_copy<<<g, 256>>>((double2*)T, (double2*)T2, nv); }, reps));
    snprintf(nm, 64, "scale-inplace g=%d", g); rep(nm, timeit([&] { k_scale<<<g, 256>>>((double2*)T, nv, 1.0); }, reps)); }
  { float ms = timeit([&] { CK(cudaMemcpyAsync(T2, T, nv * 16, cudaMemcpyDeviceToDevice)); }, reps); rep("cudaMemcpy D2D", ms); }
  for (int per : {2, 3, 4, 6, 8}) { char nm[64]; int g = sms * per;
    snprintf(nm, 64, "V0 colfixed U8 %d/SM", per); rep(nm, timeit([&] { k_v0<8><<<g, 256>>>(T, ld, R, f, p); }, reps));
    snprintf(nm, 64, "V0 colfixed U4 %d/SM", per); rep(nm, timeit([&] { k_v0<4><<<g, 256>>>(T, ld, R, f, p); }, reps));
    snprintf(nm, 64, "V0 colfixed U16 %d/SM", per); rep(nm, timeit([&] { k_v0<16><<<g, 256>>>(T, ld, R, f, p); }, reps)); }
  for (int per : {4, 8, 16}) { char nm[64]; int g = sms * per;
    snprintf(nm, 64, "V1 flat U4 %d/SM", per); rep(nm, timeit([&] { k_v1<4><<<g, 256>>>(T, ld, R, f, p); }, reps));
    snprintf(nm, 64, "V1 flat U8 %d/SM", per); rep(nm, timeit([&] { k_v1<8><<<g, 256>>>(T, ld, R, f, p); }, reps));
    snprintf(nm, 64, "V2 rowmajor U4 %d/SM", per); rep(nm, timeit([&] { k_v2<4><<<g, 256>>>(T, ld, R, f, p); }, reps));
    snprintf(nm, 64, "V2 rowmajor U8 %d/SM", per); rep(nm, timeit([&] { k_v2<8><<<g, 256>>>(T, ld, R, f, p); }, reps)); }
  for (int per : {3, 6, 8, 12, 16}) { char nm[64]; int g = sms * per; int flip = 0;
    snprintf(nm, 64, "V1m U8 fwd %d/SM", per); rep(nm, timeit([&] { k_v1m<8><<<g, 256>>>(T, ld, R, f, p, 0); }, reps));
    snprintf(nm, 64, "V1m U8 alternate-mirror %d/SM", per); rep(nm, timeit([&] { k_v1m<8><<<g, 256>>>(T, ld, R, f, p, flip); flip ^= 1; }, reps));
    snprintf(nm, 64, "V1m U4 alternate-mirror %d/SM", per); rep(nm, timeit([&] { k_v1m<4><<<g, 256>>>(T, ld, R, f, p, flip); flip ^= 1; }, reps)); }
  for (int per : {8, 16, 24}) { char nm[64]; int g = sms * per; int flip = 0;
    snprintf(nm, 64, "V1h U8 noL1 alt-mirror %d/SM", per); rep(nm, timeit([&] { k_v1h<8, 1><<<g, 256>>>(T, ld, R, f, p, flip, 0.f); flip ^= 1; }, reps));
    for (float keep : {0.15f, 0.25f, 0.35f}) { snprintf(nm, 64, "V1h U8 L2pol keep=%.2f %d/SM", keep, per); rep(nm, timeit([&] { k_v1h<8, 2><<<g, 256>>>(T, ld, R, f, p, flip, keep); flip ^= 1; }, reps)); }
    snprintf(nm, 64, "V1h U4 noL1 alt-mirror %d/SM", per); rep(nm, timeit([&] { k_v1h<4, 1><<<g, 256>>>(T, ld, R, f, p, flip, 0.f); flip ^= 1; }, reps)); }
  for (int per : {12}) { char nm[64]; int g = sms * per;
    snprintf(nm, 64, "V3 colfixed32B U4 %d/SM", per); rep(nm, timeit([&] { k_v3<4><<<g, 128>>>(T, ld64, R, f, p); }, reps));
    snprintf(nm, 64, "V3 colfixed32B U8 %d/SM", per); rep(nm, timeit([&] { k_v3<8><<<g, 128>>>(T, ld64, R, f, p); }, reps)); }
  return 0;
}
