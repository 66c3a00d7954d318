// tableau_pipelined.cu -- delayed-update primal tableau simplex with the selection of group g+1 running
// CONCURRENTLY with the tableau sweep of group g.  Bit-identical to PrimalSimplexSolver.cs:102-211.
//
// tableau_blocked.cu sweeps the tableau once per K pivots, but selection and sweep still alternate on one
// stream (select ~10 us/pivot + sweep ~11 us/pivot).  Here the sweep is OUT OF PLACE (two tableau buffers,
// sweep g reads buf[g&1] and writes buf[(g+1)&1]), so while it runs the buffer it reads is immutable and the
// selection of the next group can work from it: the entering column / leaving row of a pivot are the stale
// entries of buf[g&1] plus the K pending updates of group g plus the s pending updates of group g+1, applied
// in the original order with the reference's roundings
//     x <- (i == p_u) ? prow_u[j] : x - (f_u[i] * prow_u[j])
// => every element still sees the same operations in the same order => same bits.
//
//   stream SEL: select(0) select(1) select(2) ...      one cluster launch per group of K pivots
//   stream SW :           sweep(0)  sweep(1)  ...      select(g+1) || sweep(g)
//   select(g) waits for sweep(g-2) (its stale buffer / its group slot), sweep(g) waits for select(g).
//   The two streams live in two CUDA green contexts: 16 SMs (one GPC-resident 16-CTA cluster) for the select,
//   the other 132 SMs for the sweep -- two kernels on ordinary streams do not overlap reliably (whichever grid
//   is dispatched first fills the machine; tools/gctx_probe.cu).
//
// k_pipe_select3: one thread-block cluster (<= 16 CTAs).  Threads 0..255 of a CTA own one row each of the
//   entering column (factors of both pending groups in shared memory, RHS mirror in a register); every thread
//   owns one column (768-thread CTAs, cfg2: 16 x 768 = C-1) or up to 4 (256-thread CTAs) of the pivot row:
//   previous group's pivot-row entries in registers, this group's in shared memory, objective-row mirror in a
//   register.  Row 0 is never gathered: T[0,e] is the objective-row mirror entry and comes back as the VALUE
//   of the entering-column argmin; the RHS column is the rv mirror (the sweep does not touch it,
//   k_pipe_writeback puts it back).  The two argmins of a pivot are integer-key (value bits, index)
//   reductions: redux.sync in the warp, shared memory in the CTA, one DSMEM store per peer CTA, one cluster
//   barrier, 16 candidates to collect; the pivot element and the RHS of the winning row ride along.  Nothing is
//   stored to global memory inside the pivot loop (peers pull pivot-row / factor entries through DSMEM), so a
//   barrier's release never waits on a store; everything the sweep needs is written once at the end.
// k_pipe_sweep_ca: persistent CTAs; a CTA keeps ONE column group (its K pivot-row chunks live in registers) and
//   walks a contiguous row range; tableau rows are staged through a cp.async shared-memory ring (thread-private
//   16-byte cells, 14 rows in flight), factor records through a second ring one 32-row block ahead; two rows
//   are updated together (4 independent FP64 dependency chains per thread).
#include <cooperative_groups.h>
#include <cuda.h>  // types only: the green-context entry points are resolved through cudaGetDriverEntryPoint

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "blocked_apply.cuh"
#include "tableau.cuh"

namespace lpr {

namespace cg = cooperative_groups;

constexpr int PK = 16;            // pending pivots per group (storage and kernel instantiation)
constexpr int PMAXCTA = 16;
constexpr unsigned FULLM = 0xffffffffu;

struct PipeArgs {
  const double* Told;  // stale tableau (immutable while the previous group's sweep reads it too)
  int ld, R, C;
  const double* PRp;   // previous group: K x ld pivot rows, Rcap x PK factors, pivot rows
  const double* Fp;
  const int* pup;
  int sp;              // pending pivots of the previous group (0 for the first group of a solve)
  double* PRc;         // this group
  double* Fc;
  int* puc;
  int* count;
  double* row0;        // objective-row mirror (always current)
  double* rhs;         // RHS mirror (always current)
  int* basis;
  int* log;
  long long log_cap;
  TabState* st;
  int K;               // pivots to select in this launch
  long long* dbg;      // optional: per-phase clock64 stamps of the last launch
  long long* tl;       // optional: [select start, select end, sweep start, sweep end] globaltimer of this group
  int prefetch;        // bulk L2 prefetch of the CTA-local best row while the candidates cross the cluster
};

struct CandA {
  unsigned long long key;  // bits of the ratio (>= 0), ~0 = none
  double piv;              // entering-column value of the candidate row
  double rhs;              // RHS of the candidate row
  int idx;
  int pad;
};
struct CandB {
  unsigned long long key;  // ~bits of the (negative) objective entry, ~0 = none
  int idx;
  int pad;
};

// lexicographic (key, idx) minimum over the warp; *src = lane that holds it
__device__ __forceinline__ void warp_argmin_key(unsigned long long& key, int& idx, int& src) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned m1 = __reduce_min_sync(FULLM, hi);
  const unsigned m2 = __reduce_min_sync(FULLM, (hi == m1) ? lo : 0xffffffffu);
  const bool match = (hi == m1) && (lo == m2);
  const unsigned m3 = __reduce_min_sync(FULLM, match ? (unsigned)idx : 0xffffffffu);
  const unsigned own = __ballot_sync(FULLM, match && (unsigned)idx == m3);
  src = own ? (__ffs(own) - 1) : 0;
  key = ((unsigned long long)m1 << 32) | m2;
  idx = (int)m3;
}

__global__ void __launch_bounds__(kSelThreads) k_pipe_init(const double* T, int ld, int R, int C, double* row0,
                                                           double* rhs, int* count, TabState* st) {
  __shared__ MinIdx sm[32];
  for (int j = threadIdx.x; j < ld; j += blockDim.x) row0[j] = (j < C) ? T[j] : 0.0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) rhs[i] = TAT(T, ld, i, C - 1);
  // FindEnteringVariable :152-167
  double val = 0.0;
  int e = block_first_min(C - 1, [&](int j, double& x) { x = T[j]; return x < 0.0; }, sm, &val);
  if (threadIdx.x == 0) {
    st->enter = e;
    st->enter_val = (e >= 0) ? val : 0.0;
    st->cur = 0;
    st->do_sweep = 0;
    st->group_base = 0;
    count[0] = 0;
    count[1] = 0;
  }
}

// ---- select, third cut: more warps, no global stores inside the pivot loop ------------------------------
// ncu on k_pipe_select (2 warps per scheduler, 234 registers): issue slots 18% busy; the time goes to exposed
// fixed latencies (wait 22%), MEMBAR (16%: every cluster barrier's release waits for the global stores issued
// since the last one), barriers (15%), shared-memory / REDUX round trips (14%), instruction fetch (8%); global
// memory itself is 5%.  So: (1) NT = 768 threads per CTA when one column per thread covers the row (cfg2:
// 16 x 768 = 12288 = C-1 exactly) => 6 warps per scheduler hide each other's latencies; (2) nothing is stored
// to global memory inside the loop: the factor columns and pivot rows of the group stay in shared memory
// (peers pull what they need through DSMEM) and are written out once at the end, so a barrier's release has
// nothing to wait for; (3) candidates are reduced per CTA before they cross the cluster (16 entries to
// collect instead of 16 x warps).
constexpr int PRT = 256;  // rows owned per CTA (threads 0 .. PRT-1)

template <int NT, int NC>
__global__ void __launch_bounds__(NT, 1) k_pipe_select3(PipeArgs a) {
  constexpr int NW = NT / 32;
  constexpr int PPR = (NT > 512) ? 8 : PK;  // previous-group pivot-row entries kept in registers (80-register
                                            // budget at 768 threads); the other PK - PPR live in shared memory
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nthr = ncta * NT;
  const int gid = crank * NT + tid;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sPR = reinterpret_cast<double*>(smem_raw);   // [NC][PK][NT]  this group's pivot-row slices
  double* sFp = sPR + (size_t)NC * PK * NT;            // [PK][PRT]     previous group's factors of my row
  double* sFc = sFp + (size_t)PK * PRT;                // [PK][PRT]     this group's factors of my row
  double* sPP = sFc + (size_t)PK * PRT;                // [NC][PK - PPR][NT] previous group's pivot-row slices
  CandA* slotA = reinterpret_cast<CandA*>(sPP + (size_t)NC * (PK - PPR) * NT);  // [PMAXCTA]  one candidate per CTA
  CandB* slotB = reinterpret_cast<CandB*>(slotA + PMAXCTA);         // [PMAXCTA]
  CandA* wA = reinterpret_cast<CandA*>(slotB + PMAXCTA);            // [PRT / 32] per-warp candidates of this CTA
  CandB* wB = reinterpret_cast<CandB*>(wA + PRT / 32);              // [32]
  __shared__ double s_pe[2 * PK], s_fp[2 * PK], s_f0[PK], s_prc[PK];
  __shared__ int s_pu[2 * PK], s_el[PK];
  __shared__ unsigned s_hit;
  __shared__ double s_zobj, s_lastpiv;  // uniform scalars only thread 0 needs: kept out of the register file
  TabState* st = a.st;
  const int R = a.R, C = a.C, ld = a.ld;
  const int CW = C - 1;
  const double* T = a.Told;
  const int sp = a.sp;
  const int status = st->status;
  long long npiv = st->npiv;
  const long long npiv0 = npiv;
  const long long maxp = st->max_piv;
  int e = st->enter;
  double f0 = st->enter_val;
  if (a.tl && gid == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.tl[0] = t;
  }
  if (status != LPR_RUNNING) {  // uniform over the cluster: nobody reaches a barrier
    if (gid == 0) {
      *a.count = 0;
      st->group_base = npiv;
    }
    return;
  }
  auto stamp = [&](int q, int k) {
    if (a.dbg && gid == 0) a.dbg[q * 8 + k] = clock64();
  };
  const int i = crank * PRT + tid + 1;  // my row of the entering column (threads 0 .. PRT-1)
  const bool own_row = (tid < PRT) && (i < R);
  if (tid < 2 * PK) s_pu[tid] = (tid < sp) ? a.pup[tid] : -1;
  double rv = own_row ? a.rhs[i] : 0.0;
  bool hit_row = false;
  if (tid < PRT) {
#pragma unroll
    for (int u = 0; u < PK; u++) {
      sFp[u * PRT + tid] = (own_row && u < sp) ? a.Fp[(size_t)i * PK + u] : 0.0;
      sFc[u * PRT + tid] = 0.0;
      hit_row |= own_row && (u < sp) && (a.pup[u] == i);
    }
  }
  double r0[NC];
  double ppv[NC][PPR];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const int j = gid + c * nthr;
    r0[c] = (j < CW) ? a.row0[j] : 0.0;
#pragma unroll
    for (int u = 0; u < PK; u++) {
      const double v = (j < CW && u < sp) ? __ldg(a.PRp + (size_t)u * ld + j) : 0.0;
      if (u < PPR) ppv[c][u] = v;
      else sPP[((size_t)c * (PK - PPR) + (u - PPR)) * NT + tid] = v;
    }
  }
  if (tid == 0) {
    s_zobj = a.row0[CW];  // objective value T[0, C-1] (only CTA 0's copy is written back)
    s_lastpiv = 0.0;
  }
  __syncthreads();
  int term = LPR_RUNNING;
  int s = 0;
  for (int q = 0; q < a.K; q++, s++) {
    if (e < 0) { term = LPR_OPTIMAL; break; }
    stamp(q, 0);
    // ---- phase A -------------------------------------------------------------------------------------
    if (tid < PK) {
      s_pe[tid] = (tid < sp) ? __ldcg(a.PRp + (size_t)tid * ld + e) : 0.0;  // constant during the launch
    } else if (tid < 2 * PK) {  // this group's rows at column e: pulled from the owner's shared memory
      const int u = tid - PK;
      double v = 0.0;
      if (u < s) {
        const int ge = e % nthr, ce = e / nthr;
        v = cluster.map_shared_rank(sPR, ge / NT)[((size_t)ce * PK + u) * NT + (ge % NT)];
      }
      s_pe[tid] = v;
    } else if (tid == 2 * PK) {
      s_el[s] = e;
      s_f0[s] = f0;
    }
    double col = own_row ? TAT(T, ld, i, e) : 0.0;  // DRAM gather of the stale column
    __syncthreads();
    stamp(q, 1);
    if (w < PRT / 32) {
      unsigned long long key = ~0ull;
      int idx = INT_MAX;
      if (own_row) {
        const double* fpv = sFp + tid;
        const double* fcv = sFc + tid;
        if (!hit_row) {
          if (sp == PK) {
#pragma unroll
            for (int u = 0; u < PK; u++) col = __dsub_rn(col, __dmul_rn(fpv[u * PRT], s_pe[u]));
          } else {
            for (int u = 0; u < sp; u++) col = __dsub_rn(col, __dmul_rn(fpv[u * PRT], s_pe[u]));
          }
#pragma unroll 4
          for (int u = 0; u < s; u++) col = __dsub_rn(col, __dmul_rn(fcv[u * PRT], s_pe[PK + u]));
        } else {
          for (int u = 0; u < sp; u++)
            col = (i == s_pu[u]) ? s_pe[u] : __dsub_rn(col, __dmul_rn(fpv[u * PRT], s_pe[u]));
          for (int u = 0; u < s; u++)
            col = (i == s_pu[PK + u]) ? s_pe[PK + u] : __dsub_rn(col, __dmul_rn(fcv[u * PRT], s_pe[PK + u]));
        }
        sFc[s * PRT + tid] = col;
        // FindLeavingVariable :169-191
        if (col > 1e-9) {
          const double val = ddiv(rv, col);
          if (val >= 0.0 && val < DBL_MAX) {
            key = (val == 0.0) ? 0ull : (unsigned long long)__double_as_longlong(val);
            idx = i - 1;
          }
        }
      }
      int src;
      warp_argmin_key(key, idx, src);
      const double wpiv = __shfl_sync(FULLM, col, src);
      const double wrhs = __shfl_sync(FULLM, rv, src);
      if (lane == 0) {
        CandA c;
        c.key = key; c.piv = wpiv; c.rhs = wrhs; c.idx = idx; c.pad = 0;
        wA[w] = c;
      }
    }
    __syncthreads();
    stamp(q, 2);
    if (w == 0) {  // CTA candidate -> every CTA of the cluster
      CandA c;
      c.key = ~0ull; c.piv = 0.0; c.rhs = 0.0; c.idx = INT_MAX; c.pad = 0;
      if (lane < PRT / 32) c = wA[lane];
      int src;
      warp_argmin_key(c.key, c.idx, src);
      c.piv = __shfl_sync(FULLM, c.piv, src);
      c.rhs = __shfl_sync(FULLM, c.rhs, src);
      if (lane < ncta) cluster.map_shared_rank(slotA, lane)[crank] = c;
      if (a.prefetch && c.idx != INT_MAX && lane < 8) {
        // head start for the pivot-row read: the leaving row is one of the <= 16 CTA-local best rows; pull
        // mine towards L2 with the TMA unit (asynchronous proxy: no load/store the barrier would wait for)
        const unsigned bytes = (unsigned)(((size_t)CW * sizeof(double) + 15) & ~(size_t)15);
        const unsigned per = ((bytes / 8) + 15) & ~15u;
        const unsigned off = lane * per;
        if (off < bytes) {
          const char* rp = reinterpret_cast<const char*>(T + (size_t)(c.idx + 1) * ld) + off;
          const unsigned n = min(per, bytes - off);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rp), "r"(n) : "memory");
        }
      }
    }
    cluster.sync();
    stamp(q, 3);
    int p;
    double piv, rhsp;
    {
      CandA c;
      c.key = ~0ull; c.piv = 0.0; c.rhs = 0.0; c.idx = INT_MAX; c.pad = 0;
      if (lane < ncta) c = slotA[lane];
      int src;
      warp_argmin_key(c.key, c.idx, src);
      piv = __shfl_sync(FULLM, c.piv, src);
      rhsp = __shfl_sync(FULLM, c.rhs, src);
      p = (c.idx == INT_MAX) ? -1 : c.idx + 1;
    }
    if (p < 0) { term = LPR_UNBOUNDED; break; }
    if (maxp >= 0 && npiv >= maxp) { term = LPR_ITER_LIMIT; break; }
    stamp(q, 4);
    // ---- phase B -------------------------------------------------------------------------------------
    if (tid < PK) {
      s_fp[tid] = (tid < sp) ? __ldcg(a.Fp + (size_t)p * PK + tid) : 0.0;  // constant during the launch
    } else if (tid < 2 * PK) {  // this group's factors of row p: pulled from the owner's shared memory
      const int u = tid - PK;
      double v = 0.0;
      if (u < s) v = cluster.map_shared_rank(sFc, (p - 1) / PRT)[u * PRT + ((p - 1) % PRT)];
      s_fp[tid] = v;
    } else if (tid >= 64 && tid < 96) {
      const unsigned m = __ballot_sync(FULLM, p == s_pu[tid - 64]);
      if (tid == 64) s_hit = m;
    }
    double x[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const int j = gid + c * nthr;
      x[c] = (j < CW) ? TAT(T, ld, p, j) : 0.0;
    }
    __syncthreads();
    stamp(q, 5);
    const bool hit_p = s_hit != 0u;
    unsigned long long keyb = ~0ull;
    int idxb = INT_MAX;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const int j = gid + c * nthr;
      if (j < CW) {
        double xx = x[c];
        double* prv = sPR + (size_t)c * PK * NT + tid;
        const double* ppm = sPP + (size_t)c * (PK - PPR) * NT + tid;
        auto prev_entry = [&](int u) { return (u < PPR) ? ppv[c][u < PPR ? u : 0] : ppm[(u - PPR) * NT]; };
        if (!hit_p) {
          if (sp == PK) {
#pragma unroll
            for (int u = 0; u < PK; u++) xx = __dsub_rn(xx, __dmul_rn(s_fp[u], prev_entry(u)));
          } else {
#pragma unroll
            for (int u = 0; u < PK; u++)
              if (u < sp) xx = __dsub_rn(xx, __dmul_rn(s_fp[u], prev_entry(u)));
          }
#pragma unroll 4
          for (int u = 0; u < s; u++) xx = __dsub_rn(xx, __dmul_rn(s_fp[PK + u], prv[u * NT]));
        } else {
#pragma unroll
          for (int u = 0; u < PK; u++)
            if (u < sp) xx = (p == s_pu[u]) ? prev_entry(u) : __dsub_rn(xx, __dmul_rn(s_fp[u], prev_entry(u)));
          for (int u = 0; u < s; u++) {
            const double pv = prv[u * NT];
            xx = (p == s_pu[PK + u]) ? pv : __dsub_rn(xx, __dmul_rn(s_fp[PK + u], pv));
          }
        }
        const double pr = ddiv(xx, piv);                  // :197-199
        const double z = __dsub_rn(r0[c], __dmul_rn(f0, pr));  // :206-208 on the objective row
        if (z < 0.0) {
          const unsigned long long kb = ~(unsigned long long)__double_as_longlong(z);
          if (kb < keyb) {  // columns of one thread ascend: strict < keeps the lowest index on ties
            keyb = kb;
            idxb = j;
          }
        }
        prv[s * NT] = pr;
        r0[c] = z;
      }
    }
    const double prc = ddiv(rhsp, piv);  // normalised pivot row at the RHS column
    if (own_row) {
      rv = (i == p) ? prc : __dsub_rn(rv, __dmul_rn(col, prc));
      hit_row |= (i == p);
    }
    if (tid == 0) {
      s_pu[PK + s] = p;
      s_prc[s] = prc;
      s_zobj = __dsub_rn(s_zobj, __dmul_rn(f0, prc));
      s_lastpiv = piv;
    }
    {
      int src;
      warp_argmin_key(keyb, idxb, src);
      if (lane == 0) {
        CandB c;
        c.key = keyb; c.idx = idxb; c.pad = 0;
        wB[w] = c;
      }
    }
    __syncthreads();
    stamp(q, 6);
    if (w == 0) {
      CandB c;
      c.key = ~0ull; c.idx = INT_MAX; c.pad = 0;
      if (lane < NW) c = wB[lane];
      int src;
      warp_argmin_key(c.key, c.idx, src);
      if (lane < ncta) cluster.map_shared_rank(slotB, lane)[crank] = c;
    }
    cluster.sync();
    stamp(q, 7);
    {
      CandB c;
      c.key = ~0ull; c.idx = INT_MAX; c.pad = 0;
      if (lane < ncta) c = slotB[lane];
      int src;
      warp_argmin_key(c.key, c.idx, src);
      e = (c.idx == INT_MAX) ? -1 : c.idx;
      f0 = (c.idx == INT_MAX) ? 0.0 : __longlong_as_double((long long)~c.key);  // T[0,e] of the updated tableau
    }
    npiv++;
  }
  // ---- everything the sweep and the next launch need goes to global memory once ------------------------
  __syncthreads();
  if (own_row) {
    for (int u = 0; u < s; u++) a.Fc[(size_t)i * PK + u] = sFc[u * PRT + tid];
    a.rhs[i] = rv;
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const int j = gid + c * nthr;
    if (j < CW) {
      const double* prv = sPR + (size_t)c * PK * NT + tid;
      for (int u = 0; u < s; u++) a.PRc[(size_t)u * ld + j] = prv[u * NT];
      a.row0[j] = r0[c];
    }
  }
  if (gid == 0) {
    for (int u = 0; u < s; u++) {
      const int pu = s_pu[PK + u], eu = s_el[u];
      a.Fc[u] = s_f0[u];                       // row 0: T[0,e] at that pivot
      a.PRc[(size_t)u * ld + CW] = s_prc[u];   // RHS column of the normalised pivot row
      a.puc[u] = pu;
      if (a.log && npiv0 + u < a.log_cap) {
        a.log[2 * (npiv0 + u)] = pu;
        a.log[2 * (npiv0 + u) + 1] = eu;
      }
      if (a.basis) a.basis[pu - 1] = eu;  // :142
    }
    a.row0[CW] = s_zobj;
    *a.count = s;
    st->group_base = npiv0;
    st->npiv = npiv;
    st->enter = e;
    st->enter_val = f0;
    if (s > 0) {
      st->pivot = s_lastpiv;
      st->leave = s_pu[PK + s - 1];
    }
    if (term != LPR_RUNNING) st->status = term;
    if (a.tl) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.tl[1] = t;
    }
  }
  cluster.sync();  // no CTA exits while a peer could still pull from its shared memory
}

// ---- out-of-place sweep of one group -------------------------------------------------------------------
struct PipeSweepArgs {
  const double* src;
  double* dst;
  int ld, R;
  const double* PR;
  const double* F;
  const int* pidx;
  const int* count;
  int nch;  // 16-byte chunks per row that are swept: the columns 0 .. C-2 (RHS column = the select's rv mirror)
  long long* tl;
};

// ---- sweep with an asynchronous shared-memory ring ---------------------------------------------------------
// Same work split as k_pipe_sweep, but the tableau rows are staged through shared memory with cp.async
// (LDGSTS, 16 bytes per thread, thread-private cells => no barrier on the data ring): NSTG - B rows are in
// flight per thread ahead of the multiply/subtract chains instead of the UNROLL rows a register prefetch can
// afford.  With HBM saturated a load takes 3000-6000 clk (tools/lat_probe.cu); 2 CTAs x 12 rows x 4 KB = 96 KB
// in flight per SM covers that, 32 KB did not.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_sa(unsigned smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int KM, int B, int NSTG, int RB>
__global__ void __launch_bounds__(kSweepThreads, 2) k_pipe_sweep_ca(PipeSweepArgs a, int ncg, int nsplit) {
  static_assert(NSTG % B == 0 && RB % B == 0 && RB <= 32 && B % 2 == 0 && (NSTG & (NSTG - 1)) == 0, "ring geometry");
  constexpr unsigned kCell = sizeof(double2) * kSweepThreads;  // bytes between consecutive ring rows
  constexpr int PDB = NSTG / B - 1;  // batches in flight ahead of the one being computed
  constexpr int STG = (RB * KM / 2 + kSweepThreads - 1) / kSweepThreads;
  extern __shared__ __align__(16) unsigned char ca_smem[];
  double2* ring = reinterpret_cast<double2*>(ca_smem);                                  // [NSTG][256]
  double* sFb = reinterpret_cast<double*>(ca_smem + sizeof(double2) * NSTG * kSweepThreads);  // [2][RB*KM]
  const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring + threadIdx.x);
  __shared__ unsigned sPiv[2];
  const int s = *a.count;
  if (a.tl && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.tl[2] = t;
  }
  if (s <= 0) return;
  const int R = a.R, ld = a.ld;
  const int ldv = ld >> 1;
  const double2* S2 = reinterpret_cast<const double2*>(a.src);
  double2* D2 = reinterpret_cast<double2*>(a.dst);
  int pu[KM];
#pragma unroll
  for (int u = 0; u < KM; u++) pu[u] = (u < s) ? a.pidx[u] : -1;
  const bool full = (s == KM);
  const int tid = threadIdx.x;
  const int ntasks = ncg * nsplit;
  for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
    const int cgi = task % ncg, rs = task / ncg;
    const int r_lo = (int)((long long)R * rs / nsplit), r_hi = (int)((long long)R * (rs + 1) / nsplit);
    if (r_lo >= r_hi) continue;
    const int chunk = cgi * kSweepThreads + tid;
    const bool active = chunk < a.nch;
    const int cc = active ? chunk : 0;
    double2 pr[KM];
#pragma unroll
    for (int u = 0; u < KM; u++) pr[u] = __ldg(reinterpret_cast<const double2*>(a.PR + (size_t)u * ld) + cc);
    const int nrows = r_hi - r_lo;
    const int nblk = (nrows + RB - 1) / RB;
    const double2* gpf = S2 + cc + (size_t)r_lo * ldv;  // next row to prefetch (running pointers: the loop
    double2* gdst = D2 + cc + (size_t)r_lo * ldv;       // body is FP64-issue bound, every integer op counts)
    int pf_left = nrows;                                // rows not yet requested
    unsigned pf_cell = ring_sa;                         // shared-space address of the next ring cell to fill
    const double2* cs_cell = ring + tid;                // next ring cell to consume
    auto issue_batch = [&]() {  // the next B rows into their ring cells (one commit group, possibly empty)
#pragma unroll
      for (int k = 0; k < B; k++) {
        if (pf_left > 0) cp_async16_sa(pf_cell, gpf);
        gpf += ldv;
        pf_left--;
        pf_cell = ring_sa + ((pf_cell - ring_sa + kCell) & (NSTG * kCell - 1));
      }
      cp_async_commit();
    };
    auto stage_load = [&](int b, double2* stg) {
      const int b0 = r_lo + b * RB;
      const int n2 = (min(r_hi, b0 + RB) - b0) * KM / 2;
      const double2* f2 = reinterpret_cast<const double2*>(a.F + (size_t)b0 * KM);
#pragma unroll
      for (int k = 0; k < STG; k++) {
        const int t = tid + k * kSweepThreads;
        if (t < n2) stg[k] = __ldg(f2 + t);
      }
    };
    auto stage_store = [&](int b, const double2* stg) {
      const int b0 = r_lo + b * RB;
      const int n2 = (min(r_hi, b0 + RB) - b0) * KM / 2;
      double2* d2 = reinterpret_cast<double2*>(sFb + (size_t)(b & 1) * RB * KM);
#pragma unroll
      for (int k = 0; k < STG; k++) {
        const int t = tid + k * kSweepThreads;
        if (t < n2) d2[t] = stg[k];
      }
      if (tid < 32) {
        bool slow = !full;
        const int row = b0 + tid;
#pragma unroll
        for (int u = 0; u < KM; u++) slow |= (row == pu[u]);
        const unsigned m = __ballot_sync(FULLM, slow);
        if (tid == 0) sPiv[b & 1] = m;
      }
    };
    double2 stg[STG];
    __syncthreads();  // the previous task is done with the factor ring
#pragma unroll
    for (int bi = 0; bi < PDB; bi++) issue_batch();
    stage_load(0, stg);
    stage_store(0, stg);
    __syncthreads();
    for (int b = 0; b < nblk; b++) {
      const int b0r = b * RB, b1r = min(nrows, b0r + RB);  // rows relative to r_lo
      if (b + 1 < nblk) stage_load(b + 1, stg);
      const double2* fra = reinterpret_cast<const double2*>(sFb + (size_t)(b & 1) * RB * KM);
      unsigned pm = sPiv[b & 1];
      for (int rr = b0r; rr < b1r; rr += B) {
        issue_batch();
        cp_async_wait<PDB>();  // everything but the PDB most recent groups has landed: this batch is in the ring
        double2 x[B];
#pragma unroll
        for (int k = 0; k < B; k++) {
          x[k] = *cs_cell;
          cs_cell = ring + tid + (((cs_cell - (ring + tid)) + kSweepThreads) & (NSTG * kSweepThreads - 1));
        }
#pragma unroll
        for (int k = 0; k < B; k += 2) {
          const double2* frb = fra + KM / 2;
          const unsigned two = pm & 3u;  // CTA-uniform: pivot-row flags of rows q, q+1
          if (rr + k + 1 < b1r && two == 0u) {
            blk_apply_fast2<KM>(x[k], x[k + 1], pr, fra, frb);
            if (active) {
              gdst[0] = x[k];
              gdst[ldv] = x[k + 1];
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
              const int qq = rr + k + kk;
              if (qq < b1r) {
                const double2* fr = kk ? frb : fra;
                const bool slow = (pm >> kk) & 1u;
                const double2 y = slow ? blk_apply_gen<KM>(x[k + kk], pr, fr, r_lo + qq, s, pu)
                                       : blk_apply_fast<KM>(x[k + kk], pr, fr);
                if (active) gdst[(size_t)kk * ldv] = y;
              }
            }
          }
          fra += KM;  // two rows of KM factors = KM double2
          pm >>= 2;
          gdst += 2 * (size_t)ldv;
        }
      }
      if (b + 1 < nblk) stage_store(b + 1, stg);
      __syncthreads();
    }
    cp_async_wait<0>();
  }
  if (a.tl && threadIdx.x == 0) {  // the LAST CTA's end
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(reinterpret_cast<unsigned long long*>(a.tl + 3), (unsigned long long)t);
  }
}

// the select kernels keep the RHS column (rv / zobj mirrors) and never touch the padding: put the mirror back
// into the final tableau, keep the padding of both buffers at 0
__global__ void k_pipe_writeback(double* Tfin, double* Toth, int ld, int R, int C, const double* row0,
                                 const double* rhs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  Tfin[(size_t)i * ld + C - 1] = (i == 0) ? row0[C - 1] : rhs[i];
  for (int j = C; j < ld; j++) {
    Tfin[(size_t)i * ld + j] = 0.0;
    Toth[(size_t)i * ld + j] = 0.0;
  }
}

__global__ void k_state_reset_blk(TabState* st, long long max_piv);  // tableau_blocked.cu

// ---- SM partition: the select cluster owns 16 SMs, the sweep the other 132 (CUDA green contexts) ----------
// Two kernels on two ordinary streams do not overlap reliably: whichever grid is dispatched first fills the
// machine and a 16-CTA cluster needs 16 free SMs of one GPC at the same instant.  A green context pins each
// stream to its own SMs (tools/gctx_probe.cu: the split with MAX_POTENTIAL_CLUSTER_SIZE keeps a 16-SM group in
// which a 16-CTA cluster is schedulable).
struct GreenDev {
  int state = 0;  // 0 untried, 1 ok, -1 unavailable
  CUgreenCtx gsel = nullptr, gsw = nullptr;
  int sel_sms = 0, sw_sms = 0;
  decltype(&cuGreenCtxStreamCreate) stream_create = nullptr;
};
static GreenDev g_green[64];
static std::mutex g_green_mu;

template <class F>
static bool drv_entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

static GreenDev* green_get(int device) {
  // Green-context streams break Nsight Compute's replay (ncu dies at the first launch on them): under a profiler --
  // recognised by the injection variables its launcher exports -- the two priority streams are used instead, so a
  // profiled run still lists k_pipe_select3 / k_pipe_sweep_ca (the partitioned run is the one that is timed).
  static const bool profiled = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("CUDA_INJECTION64_PATH") ||
                               getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") || getenv("NVTX_INJECTION64_PATH");
  static const int on = getenv("LPR_PIPE_GREEN") ? atoi(getenv("LPR_PIPE_GREEN")) : (profiled ? 0 : 1);
  if (!on || device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_green_mu);
  GreenDev& G = g_green[device];
  if (G.state) return G.state > 0 ? &G : nullptr;
  G.state = -1;
  decltype(&cuDeviceGet) dev_get;
  decltype(&cuDeviceGetDevResource) get_res;
  decltype(&cuDevSmResourceSplitByCount) split;
  decltype(&cuDevResourceGenerateDesc) gen_desc;
  decltype(&cuGreenCtxCreate) create;
  if (!drv_entry("cuDeviceGet", &dev_get) || !drv_entry("cuDeviceGetDevResource", &get_res) ||
      !drv_entry("cuDevSmResourceSplitByCount", &split) || !drv_entry("cuDevResourceGenerateDesc", &gen_desc) ||
      !drv_entry("cuGreenCtxCreate", &create) || !drv_entry("cuGreenCtxStreamCreate", &G.stream_create))
    return nullptr;
  CUdevice dev;
  CUdevResource in, grp[1], rem;
  unsigned n = 1;
  if (dev_get(&dev, device) != CUDA_SUCCESS) return nullptr;
  if (get_res(dev, &in, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if (in.sm.smCount < 4 * PMAXCTA) return nullptr;
  if (split(grp, &n, &in, &rem, CU_DEV_SM_RESOURCE_SPLIT_MAX_POTENTIAL_CLUSTER_SIZE, PMAXCTA) != CUDA_SUCCESS || n < 1)
    return nullptr;
  if (grp[0].sm.smCount < (unsigned)PMAXCTA || rem.sm.smCount < 1) return nullptr;
  CUdevResourceDesc da, db;
  if (gen_desc(&da, &grp[0], 1) != CUDA_SUCCESS || gen_desc(&db, &rem, 1) != CUDA_SUCCESS) return nullptr;
  if (create(&G.gsel, da, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  if (create(&G.gsw, db, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  G.sel_sms = (int)grp[0].sm.smCount;
  G.sw_sms = (int)rem.sm.smCount;
  G.state = 1;
  return &G;
}

void tab_pipe_free(lpr_tab* h) {
  auto& P = h->pipe;
  for (int k = 0; k < 2; k++) {
    cudaFree(P.pr[k]);
    cudaFree(P.f[k]);
    if (P.ev_sel[k]) cudaEventDestroy(P.ev_sel[k]);
    if (P.ev_sw[k]) cudaEventDestroy(P.ev_sw[k]);
    P.pr[k] = P.f[k] = nullptr;
    P.ev_sel[k] = P.ev_sw[k] = nullptr;
  }
  cudaFree(P.pidx);
  cudaFree(P.count);
  cudaFree(P.row0);
  cudaFree(P.rhs);
  if (P.ev_in) cudaEventDestroy(P.ev_in);
  if (P.ev_out) cudaEventDestroy(P.ev_out);
  if (P.s_sel) cudaStreamDestroy(P.s_sel);
  if (P.s_sw) cudaStreamDestroy(P.s_sw);
  P = lpr_tab::PipeRes();
}

static int pipe_alloc(lpr_tab* h) {
  auto& P = h->pipe;
  if (P.k == PK) return LPR_OK;
  for (int k = 0; k < 2; k++) {
    LPR_CUDA(cudaMalloc(&P.pr[k], sizeof(double) * (size_t)PK * h->ld));
    LPR_CUDA(cudaMalloc(&P.f[k], sizeof(double) * (size_t)PK * h->Rcap));
    LPR_CUDA(cudaMemset(P.pr[k], 0, sizeof(double) * (size_t)PK * h->ld));
    LPR_CUDA(cudaMemset(P.f[k], 0, sizeof(double) * (size_t)PK * h->Rcap));
    LPR_CUDA(cudaEventCreateWithFlags(&P.ev_sel[k], cudaEventDisableTiming));
    LPR_CUDA(cudaEventCreateWithFlags(&P.ev_sw[k], cudaEventDisableTiming));
  }
  LPR_CUDA(cudaMalloc(&P.pidx, sizeof(int) * 2 * PK));
  LPR_CUDA(cudaMemset(P.pidx, 0xff, sizeof(int) * 2 * PK));
  LPR_CUDA(cudaMalloc(&P.count, sizeof(int) * 2));
  LPR_CUDA(cudaMemset(P.count, 0, sizeof(int) * 2));
  LPR_CUDA(cudaMalloc(&P.row0, sizeof(double) * h->ld));
  LPR_CUDA(cudaMalloc(&P.rhs, sizeof(double) * h->Rcap));
  LPR_CUDA(cudaEventCreateWithFlags(&P.ev_in, cudaEventDisableTiming));
  LPR_CUDA(cudaEventCreateWithFlags(&P.ev_out, cudaEventDisableTiming));
  GreenDev* G = green_get(h->device);
  if (G) {
    CUstream a = nullptr, b = nullptr;
    if (G->stream_create(&a, G->gsel, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
        G->stream_create(&b, G->gsw, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS) {
      P.s_sel = (cudaStream_t)a;
      P.s_sw = (cudaStream_t)b;
      P.sw_sms = G->sw_sms;
    } else {
      if (a) cudaStreamDestroy((cudaStream_t)a);
      if (b) cudaStreamDestroy((cudaStream_t)b);
    }
  }
  if (!P.s_sel) {  // no SM partition: two priority streams (the kernels overlap only when the scheduler lets them)
    int lo = 0, hi = 0;  // numerically lower = higher priority
    LPR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LPR_CUDA(cudaStreamCreateWithPriority(&P.s_sel, cudaStreamNonBlocking, hi));
    LPR_CUDA(cudaStreamCreateWithPriority(&P.s_sw, cudaStreamNonBlocking, lo));
    P.sw_sms = 0;
  }
  P.k = PK;
  return LPR_OK;
}

static size_t ca_smem_bytes(int nstg) {
  return sizeof(double2) * (size_t)nstg * kSweepThreads + sizeof(double) * 2 * 32 * PK;
}

// select v3 geometry: smallest cluster whose CTAs (256 row owners each) cover the rows; one column per thread
// with 768-thread CTAs when that covers the row, else up to 4 columns per thread with 256-thread CTAs
static void pipe_geometry3(const lpr_tab* h, int* ncta, int* nt, int* nc) {
  *ncta = 0;
  *nt = 0;
  *nc = 0;
  const int rows = h->R - 1, cw = h->C - 1;
  for (int c = 1; c <= PMAXCTA; c <<= 1) {
    if (rows > c * PRT) continue;
    if (cw <= c * 256) { *ncta = c; *nt = 256; *nc = 1; return; }
    if (cw <= c * 768) { *ncta = c; *nt = 768; *nc = 1; return; }
    if (cw <= c * 256 * 4) { *ncta = c; *nt = 256; *nc = 4; return; }  // (3 c 256, 4 c 256]: four columns per thread
  }
}
static size_t pipe_select3_smem(int nt, int nc) {
  const int ppr = nt > 512 ? 8 : PK;  // as in the kernel
  return sizeof(double) * ((size_t)nc * PK * nt + 2 * (size_t)PK * PRT + (size_t)nc * (PK - ppr) * nt) +
         sizeof(CandA) * (PMAXCTA + PRT / 32) + sizeof(CandB) * (PMAXCTA + 32);
}

using SelectFn = void (*)(PipeArgs);
static SelectFn pipe_select3_fn(int nt, int nc) {
  if (nt == 768) return k_pipe_select3<768, 1>;
  return nc == 1 ? k_pipe_select3<256, 1> : k_pipe_select3<256, 4>;  // the geometry only yields 1 or 4
}
static bool pipe_prepare3(int ncta, int nt, int nc) {
  static int ok[2][5][PMAXCTA + 1] = {};
  int& slot = ok[nt == 768][nc][ncta];
  if (slot) return slot > 0;
  SelectFn fn = pipe_select3_fn(nt, nc);
  const size_t smem = pipe_select3_smem(nt, nc);
  bool good = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
  if (good && ncta > 8)
    good = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (good) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(nt);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    good = cudaOccupancyMaxActiveClusters(&n, fn, &cfg) == cudaSuccess && n >= 1;
  }
  if (!good) cudaGetLastError();
  slot = good ? 1 : -1;
  return good;
}

bool tab_pipe_applicable(const lpr_tab* h) {
  static const int on = getenv("LPR_TAB_PIPE") ? atoi(getenv("LPR_TAB_PIPE")) : 1;
  if (!on) return false;
  int ncta, nc, nt;
  pipe_geometry3(h, &ncta, &nt, &nc);
  return ncta && pipe_prepare3(ncta, nt, nc);
}

// Solve() for LPR_RULE_PRIMAL, K delayed pivots per sweep, selection overlapped with the previous sweep.
// Same contract as tab_solve_blocked.
int tab_solve_pipelined(lpr_tab* h, int K, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log,
                        int64_t log_cap, bool time_sweeps) {
  K = std::max(2, std::min(K, PK));
  int rc = tab_ensure_T2(h);
  if (rc) return rc;
  if ((rc = pipe_alloc(h))) return rc;
  if (pivot_log && log_cap > 0) {
    long long want = log_cap;
    if (max_pivots >= 0) want = std::min<long long>(want, max_pivots + 1);
    want = std::min<long long>(want, 1LL << 24);
    if ((rc = tab_ensure_log(h, want))) return rc;
  }
  auto& P = h->pipe;
  int ncta, nc, nt;
  pipe_geometry3(h, &ncta, &nt, &nc);
  SelectFn fn = pipe_select3_fn(nt, nc);
  const size_t smem = pipe_select3_smem(nt, nc);
  double* buf[2] = {h->T, h->T2};

  long long* d_dbg = nullptr;
  if (getenv("LPR_BLK_TIMING")) {
    LPR_CUDA(cudaMalloc(&d_dbg, sizeof(long long) * 8 * PK));
    LPR_CUDA(cudaMemset(d_dbg, 0, sizeof(long long) * 8 * PK));
  }
  static const int prefetch_on = getenv("LPR_PIPE_PREFETCH") ? atoi(getenv("LPR_PIPE_PREFETCH")) : 0;
  long long* d_tl = nullptr;
  const int TLG = 24;
  if (getenv("LPR_PIPE_TIMELINE")) {
    LPR_CUDA(cudaMalloc(&d_tl, sizeof(long long) * 4 * TLG));
    LPR_CUDA(cudaMemset(d_tl, 0, sizeof(long long) * 4 * TLG));
  }
  // sweep geometry: persistent CTAs, (column group, row split) tasks; leave the select cluster its slots
  static const int overlap = getenv("LPR_PIPE_OVERLAP") ? atoi(getenv("LPR_PIPE_OVERLAP")) : 1;
  static const int reserve_env = getenv("LPR_PIPE_RESERVE") ? atoi(getenv("LPR_PIPE_RESERVE")) : -1;
  static const int sweep_kind = getenv("LPR_PIPE_SWEEP") ? atoi(getenv("LPR_PIPE_SWEEP")) : 2;
  static int resident = 0;
  if (!resident) {
    cudaFuncSetAttribute(k_pipe_sweep_ca<PK, 4, 16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ca_smem_bytes(16));
    cudaFuncSetAttribute(k_pipe_sweep_ca<PK, 2, 16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ca_smem_bytes(16));
    cudaFuncSetAttribute(k_pipe_sweep_ca<PK, 4, 8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ca_smem_bytes(8));
    cudaFuncSetAttribute(k_pipe_sweep_ca<PK, 2, 8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ca_smem_bytes(8));
    int r = 0;
    if (sweep_kind == 1)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_pipe_sweep_ca<PK, 4, 16, 32>, kSweepThreads, ca_smem_bytes(16));
    else if (sweep_kind >= 3)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_pipe_sweep_ca<PK, 4, 8, 32>, kSweepThreads, ca_smem_bytes(8));
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_pipe_sweep_ca<PK, 2, 16, 32>, kSweepThreads, ca_smem_bytes(16));
    resident = std::max(1, r);
  }
  const int nch = (h->C - 1 + 1) / 2;  // chunks that hold the columns 0 .. C-2
  const int ncg = std::max(1, (nch + kSweepThreads - 1) / kSweepThreads);
  const int reserve = reserve_env >= 0 ? reserve_env : ((overlap && !P.sw_sms) ? ncta : 0);
  const int slots = std::max(1, (P.sw_sms ? P.sw_sms : h->sms) * resident - reserve);
  static const int split_env = getenv("LPR_PIPE_SPLIT") ? atoi(getenv("LPR_PIPE_SPLIT")) : 0;
  int nsplit = split_env > 0 ? split_env : std::max(1, slots / ncg);  // > slots/ncg: short CTAs, several waves
  nsplit = std::min(nsplit, std::max(1, (h->R + 31) / 32));
  const int gs = split_env > 0 ? ncg * nsplit : std::min(ncg * nsplit, slots);

  std::vector<cudaEvent_t> sw_ev;
  h->last_sweep_us = 0.f;
  LPR_CUDA(cudaEventRecord(P.ev_in, h->stream));
  LPR_CUDA(cudaStreamWaitEvent(P.s_sel, P.ev_in, 0));
  LPR_CUDA(cudaStreamWaitEvent(P.s_sw, P.ev_in, 0));
  LPR_CUDA(cudaEventRecord(h->ev0, P.s_sel));
  k_state_reset_blk<<<1, 1, 0, P.s_sel>>>(h->st, (long long)max_pivots);
  LPR_LAUNCH_CHECK();
  k_pipe_init<<<1, kSelThreads, 0, P.s_sel>>>(h->T, h->ld, h->R, h->C, P.row0, P.rhs, P.count, h->st);
  LPR_LAUNCH_CHECK();

  auto launch_group = [&](long long g) -> int {
    const int sl = (int)(g & 1);
    if (g >= 2) LPR_CUDA(cudaStreamWaitEvent(P.s_sel, P.ev_sw[sl], 0));  // sweep(g-2): my stale buffer, my slot
    if (!overlap && g >= 1) LPR_CUDA(cudaStreamWaitEvent(P.s_sel, P.ev_sw[sl ^ 1], 0));
    PipeArgs a;
    a.Told = (g == 0) ? buf[0] : buf[(g + 1) & 1];
    a.ld = h->ld; a.R = h->R; a.C = h->C;
    a.PRp = P.pr[sl ^ 1]; a.Fp = P.f[sl ^ 1]; a.pup = P.pidx + (sl ^ 1) * PK;
    a.sp = (g == 0) ? 0 : K;
    a.PRc = P.pr[sl]; a.Fc = P.f[sl]; a.puc = P.pidx + sl * PK; a.count = P.count + sl;
    a.row0 = P.row0; a.rhs = P.rhs;
    a.basis = h->basis;
    a.log = (pivot_log && log_cap > 0) ? h->log : nullptr;
    a.log_cap = (pivot_log && log_cap > 0) ? h->log_cap : 0;
    a.st = h->st;
    a.K = K;
    a.dbg = d_dbg;
    a.tl = (d_tl && g < TLG) ? d_tl + 4 * g : nullptr;
    a.prefetch = prefetch_on;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(nt);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = P.s_sel;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t ce = cudaLaunchKernelEx(&cfg, fn, a);
    if (ce != cudaSuccess)
      return fail(LPR_E_CUDA, "pipelined select launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    count_launch();
    LPR_CUDA(cudaEventRecord(P.ev_sel[sl], P.s_sel));
    LPR_CUDA(cudaStreamWaitEvent(P.s_sw, P.ev_sel[sl], 0));
    const bool timed = time_sweeps && sw_ev.size() < 256;
    if (timed) {
      cudaEvent_t e0, e1;
      LPR_CUDA(cudaEventCreate(&e0));
      LPR_CUDA(cudaEventCreate(&e1));
      sw_ev.push_back(e0);
      sw_ev.push_back(e1);
      LPR_CUDA(cudaEventRecord(e0, P.s_sw));
    }
    PipeSweepArgs w;
    w.src = buf[g & 1];
    w.dst = buf[(g + 1) & 1];
    w.ld = h->ld; w.R = h->R;
    w.PR = P.pr[sl]; w.F = P.f[sl]; w.pidx = P.pidx + sl * PK; w.count = P.count + sl;
    w.nch = nch;
    w.tl = (d_tl && g < TLG) ? d_tl + 4 * g : nullptr;
    if (sweep_kind == 1)
      k_pipe_sweep_ca<PK, 4, 16, 32><<<gs, kSweepThreads, ca_smem_bytes(16), P.s_sw>>>(w, ncg, nsplit);
    else if (sweep_kind == 4)
      k_pipe_sweep_ca<PK, 4, 8, 32><<<gs, kSweepThreads, ca_smem_bytes(8), P.s_sw>>>(w, ncg, nsplit);
    else if (sweep_kind == 5)
      k_pipe_sweep_ca<PK, 2, 8, 32><<<gs, kSweepThreads, ca_smem_bytes(8), P.s_sw>>>(w, ncg, nsplit);
    else
      k_pipe_sweep_ca<PK, 2, 16, 32><<<gs, kSweepThreads, ca_smem_bytes(16), P.s_sw>>>(w, ncg, nsplit);
    if (cudaGetLastError() != cudaSuccess)
      return fail(LPR_E_CUDA, "pipelined sweep launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    count_launch();
    if (timed) LPR_CUDA(cudaEventRecord(sw_ev.back(), P.s_sw));
    LPR_CUDA(cudaEventRecord(P.ev_sw[sl], P.s_sw));
    return LPR_OK;
  };

  static const int groups_per_batch = std::max(1, getenv("LPR_TAB_BATCH") ? atoi(getenv("LPR_TAB_BATCH")) / 4 : 8);
  long long g = 0;
  int slot = 0, pending = 0, ngroups = 1;
  while (true) {
    for (int k = 0; k < ngroups; k++, g++)
      if ((rc = launch_group(g))) return rc;
    // status poll on the handle's own stream: a D2H copy queued between two selects stalls the select stream
    // for ~20 us (measured), an event dependency does not
    LPR_CUDA(cudaEventRecord(P.ev_out, P.s_sel));
    LPR_CUDA(cudaStreamWaitEvent(h->stream, P.ev_out, 0));
    LPR_CUDA(cudaMemcpyAsync(&h->st_host[slot], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaEventRecord(h->evb[slot], h->stream));
    pending++;
    if (pending == 2 || ngroups < groups_per_batch) {
      const int old = (pending == 2) ? (slot ^ 1) : slot;
      LPR_CUDA(cudaEventSynchronize(h->evb[old]));
      pending--;
      if (h->st_host[old].status != LPR_RUNNING) break;
    }
    slot ^= 1;
    ngroups = std::min(groups_per_batch, ngroups * 2);
  }
  // drain: every sweep done, then stop the clock on the select stream
  LPR_CUDA(cudaStreamWaitEvent(P.s_sel, P.ev_sw[0], 0));
  LPR_CUDA(cudaStreamWaitEvent(P.s_sel, P.ev_sw[1], 0));
  LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, P.s_sel));
  LPR_CUDA(cudaStreamSynchronize(P.s_sel));
  const long long npiv = h->st_host[0].npiv;
  // the current tableau is in buf[#effective sweeps & 1]; its RHS column is the select kernels' mirror
  const long long neff = (npiv + K - 1) / K;
  if (neff & 1) std::swap(h->T, h->T2);
  k_pipe_writeback<<<(h->R + 255) / 256, 256, 0, P.s_sel>>>(h->T, h->T2, h->ld, h->R, h->C, P.row0, P.rhs);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaEventRecord(h->ev1, P.s_sel));
  LPR_CUDA(cudaStreamSynchronize(P.s_sel));
  LPR_CUDA(cudaStreamSynchronize(P.s_sw));
  LPR_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  if (d_dbg) {
    long long hd[8 * PK];
    cudaMemcpy(hd, d_dbg, sizeof hd, cudaMemcpyDeviceToHost);
    for (int q = 0; q < K && hd[q * 8]; q++) {
      fprintf(stderr, "[pipe timing] pivot %2d:", q);
      for (int k = 1; k < 8; k++) fprintf(stderr, " %5lld", hd[q * 8 + k] - hd[q * 8 + k - 1]);
      if (q + 1 < K && hd[(q + 1) * 8]) fprintf(stderr, " | next %5lld", hd[(q + 1) * 8] - hd[q * 8 + 7]);
      fprintf(stderr, "  clk (stageA chainA publishSyncA collectA stageB chainB publishSyncB)\n");
    }
    cudaFree(d_dbg);
  }
  if (d_tl) {
    long long ht[4 * 24];
    cudaMemcpy(ht, d_tl, sizeof ht, cudaMemcpyDeviceToHost);
    const long long t0 = ht[0];
    for (int q = 0; q < TLG && ht[4 * q]; q++)
      fprintf(stderr, "[pipe timeline] group %2d: select %8.1f .. %8.1f us   sweep %8.1f .. %8.1f us\n", q,
              (ht[4 * q] - t0) * 1e-3, (ht[4 * q + 1] - t0) * 1e-3, (ht[4 * q + 2] - t0) * 1e-3,
              (ht[4 * q + 3] - t0) * 1e-3);
    cudaFree(d_tl);
  }
  if (time_sweeps) {
    double sum = 0.0;
    int cnt = 0;
    const int real = (int)std::min<long long>(neff, (long long)sw_ev.size() / 2);
    for (int k = std::min(2, real / 2); k < real; k++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sw_ev[2 * k], sw_ev[2 * k + 1]) == cudaSuccess) { sum += ms; cnt++; }
    }
    if (cnt) h->last_sweep_us = (float)(sum * 1e3 / cnt);
    for (auto ev : sw_ev) cudaEventDestroy(ev);
  }
  if (status) *status = h->st_host[0].status;
  if (n_pivots) *n_pivots = npiv;
  if (pivot_log && log_cap > 0 && npiv > 0) {
    long long cnt = std::min<long long>(std::min<long long>(npiv, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(pivot_log, h->log, sizeof(int) * 2 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  return LPR_OK;
}

}  // namespace lpr
