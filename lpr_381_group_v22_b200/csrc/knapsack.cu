// knapsack.cu -- branch & bound knapsack (Program.cs:430-471) and its DP arbiter on the device.
//
// The reference's KnapsackBranchBoundSimplex / KnapsackBranchBoundSolver bodies are missing
// (IntegerProgramming/KnapsackBranchBoundSolver.cs:9-11 is an empty class); the call site is the only
// contract.  Specification implemented here (and by oracle/lpr_oracle.cpp orc_knap_bb):
//   * items ranked by value/weight descending, ties by lower original id;
//   * a node fixes some ranked positions to 0/1; its relaxation takes the fixed-1 items, then fills
//     the free items greedily in rank order and stops at the first one that does not fit (critical
//     item k): bound = value + v_k * (cap_left / w_k); no critical item or cap_left == 0 => candidate;
//   * children: x_k = 0 (DFS-first) then x_k = 1; nodes with bound <= incumbent are fathomed;
//   * the answer is the DFS-first optimal candidate.  Nodes carry their DFS path as a bit string so
//     that (value, path) selects the same incumbent for ANY exploration order / GPU count.
// Device layout: the open-node pool is an array of fixed-size records in HBM
//   [ fixmask : W words | fixval : W words | key : W words | depth | (k, side) | cap_k | val_k ]   W = ceil(n/64)
// (3.8 KB at n = 10^4).  The last three words are the greedy state inherited from the parent, whose
// critical item was k: cap_k / val_k = capacity left / value collected when the parent's fill stopped at k.  A
// child differs from its parent in that one item, so its relaxation is a short walk from k instead of two passes
// over all n items (k_knap_eval, one thread per node):
//   * x_k = 0: the items before k stay as they were; the fill resumes at k+1 with (cap_k, val_k);
//   * x_k = 1: k is forced in (cap_k - w_k < 0), so free items are given back from k-1 downwards until the
//     rest fits; the last one given back is the new critical item.
// Sums are therefore formed in a different order than a front-to-back scan; they are exact (and the result
// identical) for integer-valued weights and values below 2^53, which is what the call site passes
// (Program.cs:444-448) and what the specification above assumes.  A second kernel writes the two children of the
// surviving nodes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

namespace lpr {

struct KnapEval {
  double val;   // candidate value or bound
  double cap;   // capacity left when the fill stopped (state handed to the children)
  double base;  // value collected when the fill stopped
  int crit;     // critical ranked position, -1 = none
  int type;     // 0 infeasible, 1 candidate, 2 branch
  int cmp;      // DFS order of the node's key against the incumbent's (-1 before, 0 same, 1 after) when val ties with
                // the incumbent value the kernel was given; kCmpNone otherwise
  int pad;
};
constexpr int kCmpNone = 3;

// DFS order of two path keys (bit strings): first differing bit decides, a prefix precedes its extensions
__host__ __device__ inline int knap_key_order(const uint64_t* a, int abits, const uint64_t* b, int bbits) {
  const int n = abits < bbits ? abits : bbits;
  for (int w0 = 0; w0 * 64 < n; w0++) {
    uint64_t x = a[w0] ^ b[w0];
    const int left = n - w0 * 64;
    if (left < 64) x &= (1ull << left) - 1ull;
    if (x) {
      const uint64_t low = x & (~x + 1ull);  // lowest differing bit = earliest branch
      return (a[w0] & low) ? 1 : -1;
    }
  }
  if (abits == bbits) return 0;
  return abits < bbits ? -1 : 1;
}

constexpr int kRecTail = 4;  // depth | (k, side) | cap_k | val_k
__host__ __device__ inline uint64_t knap_pack_state(int k, int side) {
  return (uint64_t)(uint32_t)k | ((uint64_t)(uint32_t)side << 32);
}

// Device-resident control block: the tree is walked level by level WITHOUT the host in the loop.  One level, six
// launches back to back, all but two of them spread over the whole batch:
//   k_knap_eval   one thread per node of the batch (top ev_nb records of the stack, in place): relaxation, and the
//                 CTA's best candidate (value, then DFS-first key)
//   k_knap_inc    one small CTA: best candidate of the level -> incumbent (strict improvement, or a tie with an
//                 earlier key); the incumbent's record stays on the device
//   k_knap_flag   per node: does it survive the (new) incumbent?  one ballot word per warp, one count per CTA
//   k_knap_scan   one small CTA: exclusive scan of the CTA counts, stack bookkeeping, node budget, next batch window
//   k_knap_gather surviving parents -> staging (children overwrite the batch's slots); the j-th survivor is found from
//                 the scanned counts and the ballot words, so survivors stay in batch order (the stack order, hence
//                 the node count, does not depend on scheduling)
//   k_knap_expand both children of every survivor
// Round 2's first cut did incumbent, pruning and compaction in ONE 1024-thread CTA (k_knap_plan): 75 us per 16384-node
// level against 12 us for the evaluation (ncu: long-scoreboard stalls 95 per issue) -- the level loop ran at the
// speed of that CTA.  The host enqueues several levels back to back and only then reads this block (lpr_knap_run).
struct KnapCtl {
  long long open;       // stack size (records)
  long long processed;  // nodes evaluated since creation
  long long max_nodes;  // absolute budget on `processed` for the current run, < 0 = none
  long long pool_cap;
  long long ev_first;   // batch of the next k_knap_eval: records [ev_first, ev_first + ev_nb)
  long long ex_first;   // children of the current level are written from here (= first slot of its batch)
  int ev_nb, ex_nj;
  int has_inc, inc_bits, inc_crit, inc_version;
  double inc_val;
  int error, stop, levels, batch;
  int ex_ncta, reserved;  // evaluation CTAs of the level being expanded (extent of the scanned counts)
};

constexpr int kEvT = 128;  // threads per CTA of the per-node kernels (eval, flag, gather-by-node)
struct KnapCand {          // best candidate of one evaluation CTA
  double val;
  int idx;  // node index inside the batch, -1 = none
  int pad;
};

// is candidate node a (value va) ahead of candidate node b?  value first, then the DFS-first key
__device__ __forceinline__ bool knap_cand_ahead(const uint64_t* pool, size_t rec_words, int W, long long first, double va,
                                                int a, double vb, int b) {
  if (b < 0) return a >= 0;
  if (a < 0) return false;
  if (va != vb) return va > vb;
  const uint64_t* ra = pool + (size_t)(first + a) * rec_words;
  const uint64_t* rb = pool + (size_t)(first + b) * rec_words;
  return knap_key_order(ra + 2 * (size_t)W, (int)ra[3 * (size_t)W], rb + 2 * (size_t)W, (int)rb[3 * (size_t)W]) < 0;
}

// relaxation of one node: a walk from the parent's critical item (see the header)
__device__ __forceinline__ KnapEval knap_eval_node(const uint64_t* __restrict__ rec, int W, int n,
                                                   const double* __restrict__ w, const double* __restrict__ v) {
  const uint64_t st = rec[3 * (size_t)W + 1];
  const int k = (int)(uint32_t)st, side = (int)(st >> 32);
  double cap = __longlong_as_double((long long)rec[3 * (size_t)W + 2]);
  double val = __longlong_as_double((long long)rec[3 * (size_t)W + 3]);
  // free (not yet branched on) positions are found a 64-bit word of the fix mask at a time: deep nodes have hundreds of
  // fixed items around their critical item, and testing them one by one was most of the walk
  auto prev_free = [&](int p) {  // highest free position <= p, -1 if none
    if (p < 0) return -1;
    int wi = p >> 6;
    uint64_t m = ~rec[wi] & (((p & 63) == 63) ? ~0ull : ((2ull << (p & 63)) - 1ull));
    while (true) {
      if (m) return wi * 64 + 63 - __clzll((long long)m);
      if (--wi < 0) return -1;
      m = ~rec[wi];
    }
  };
  auto next_free = [&](int p) {  // lowest free position >= p, n if none
    if (p >= n) return n;
    int wi = p >> 6;
    uint64_t m = ~rec[wi] & (~0ull << (p & 63));
    while (true) {
      if (m) {
        const int q = wi * 64 + __ffsll((long long)m) - 1;
        return q < n ? q : n;
      }
      if (++wi >= W) return n;
      m = ~rec[wi];
    }
  };
  int crit = -1;
  bool infeasible = k < 0 && cap < 0.0;  // negative capacity at the root
  if (infeasible) {
  } else if (side == 1 && k >= 0) {
    cap = __dsub_rn(cap, w[k]);  // < 0: k did not fit in the parent
    val = __dadd_rn(val, v[k]);
    int p = prev_free(k - 1);
    while (p >= 0) {
      cap = __dadd_rn(cap, w[p]);
      val = __dsub_rn(val, v[p]);
      if (cap >= 0.0) break;
      p = prev_free(p - 1);
    }
    if (p < 0) infeasible = true;  // the fixed-1 items alone exceed the capacity
    crit = p;
  } else {
    int p = next_free(k + 1);  // k = -1 for the root
    while (p < n) {
      const double wp = w[p];
      if (wp <= cap) {
        cap = __dsub_rn(cap, wp);
        val = __dadd_rn(val, v[p]);
        p = next_free(p + 1);
      } else {
        crit = p;
        break;
      }
    }
  }
  KnapEval ev;
  ev.cap = cap;
  ev.base = val;
  ev.crit = crit;
  if (infeasible) {
    ev.val = 0.0;
    ev.crit = -1;
    ev.type = 0;
  } else if (crit < 0 || cap == 0.0) {
    ev.val = val;
    ev.type = 1;
  } else {
    ev.val = __dadd_rn(val, __dmul_rn(v[crit], __ddiv_rn(cap, w[crit])));
    ev.type = 2;
  }
  ev.cmp = kCmpNone;
  ev.pad = 0;
  return ev;
}

// one thread per node, then the CTA's best candidate
__global__ void __launch_bounds__(kEvT) k_knap_eval(const KnapCtl* __restrict__ ctl, const uint64_t* __restrict__ pool,
                                                    size_t rec_words, int W, int n, const double* __restrict__ w,
                                                    const double* __restrict__ v, KnapEval* __restrict__ out,
                                                    KnapCand* __restrict__ cands) {
  __shared__ double s_val[kEvT / 32];
  __shared__ int s_idx[kEvT / 32];
  const int nb = ctl->ev_nb;
  const long long first = ctl->ev_first;
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  double cval = 0.0;
  int cidx = -1;
  if (node < nb) {
    const KnapEval ev = knap_eval_node(pool + (size_t)(first + node) * rec_words, W, n, w, v);
    out[node] = ev;
    if (ev.type == 1) {
      cval = ev.val;
      cidx = node;
    }
  }
  // best candidate of this CTA: warp shuffle reduction with the (value, key) order, then across the warps
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, cval, o);
    const int oi = __shfl_xor_sync(0xffffffffu, cidx, o);
    if (knap_cand_ahead(pool, rec_words, W, first, ov, oi, cval, cidx)) {
      cval = ov;
      cidx = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_val[threadIdx.x >> 5] = cval;
    s_idx[threadIdx.x >> 5] = cidx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < kEvT / 32; q++)
      if (knap_cand_ahead(pool, rec_words, W, first, s_val[q], s_idx[q], cval, cidx)) {
        cval = s_val[q];
        cidx = s_idx[q];
      }
    KnapCand c;
    c.val = cval;
    c.idx = cidx;
    c.pad = 0;
    cands[blockIdx.x] = c;
  }
}

// best candidate of the level (over the evaluation CTAs) -> incumbent
constexpr int kIncT = 256;
__global__ void __launch_bounds__(kIncT) k_knap_inc(KnapCtl* ctl, const uint64_t* __restrict__ pool, size_t rec_words, int W,
                                                    const KnapEval* __restrict__ evals, const KnapCand* __restrict__ cands,
                                                    uint64_t* inc_key, uint64_t* inc_rec) {
  __shared__ double s_val[kIncT];
  __shared__ int s_idx[kIncT];
  __shared__ int s_upd;
  const int tid = threadIdx.x;
  const int nb = ctl->ev_nb;
  if (nb <= 0 || ctl->stop) return;
  const long long first = ctl->ev_first;
  const int ncta = (nb + kEvT - 1) / kEvT;
  double cval = 0.0;
  int cidx = -1;
  for (int q = tid; q < ncta; q += kIncT) {
    const KnapCand c = cands[q];
    if (knap_cand_ahead(pool, rec_words, W, first, c.val, c.idx, cval, cidx)) {
      cval = c.val;
      cidx = c.idx;
    }
  }
  s_val[tid] = cval;
  s_idx[tid] = cidx;
  __syncthreads();
  for (int o = kIncT / 2; o > 0; o >>= 1) {
    if (tid < o && knap_cand_ahead(pool, rec_words, W, first, s_val[tid + o], s_idx[tid + o], s_val[tid], s_idx[tid])) {
      s_val[tid] = s_val[tid + o];
      s_idx[tid] = s_idx[tid + o];
    }
    __syncthreads();
  }
  const int best = s_idx[0];
  if (tid == 0) {
    int upd = 0;
    if (best >= 0) {
      const uint64_t* rb = pool + (size_t)(first + best) * rec_words;
      const double bv = s_val[0];
      if (!ctl->has_inc || bv > ctl->inc_val ||
          (bv == ctl->inc_val && knap_key_order(rb + 2 * (size_t)W, (int)rb[3 * (size_t)W], inc_key, ctl->inc_bits) < 0))
        upd = 1;
    }
    s_upd = upd;
  }
  __syncthreads();
  if (!s_upd) return;
  // the incumbent's record stays on the device; the host fetches it when asked (lpr_knap_get_incumbent)
  const uint64_t* src = pool + (size_t)(first + best) * rec_words;
  for (int t = tid; t < (int)rec_words; t += kIncT) inc_rec[t] = src[t];
  for (int t = tid; t < W; t += kIncT) inc_key[t] = src[2 * (size_t)W + t];
  __syncthreads();
  if (tid == 0) {
    ctl->has_inc = 1;
    ctl->inc_val = s_val[0];
    ctl->inc_bits = (int)src[3 * (size_t)W];
    ctl->inc_crit = evals[best].crit;
    ctl->inc_version++;
  }
}

// a branch node survives when its bound beats the incumbent, or ties with it while preceding it in DFS order:
// one ballot word per warp, one count per CTA
__global__ void __launch_bounds__(kEvT) k_knap_flag(const KnapCtl* __restrict__ ctl, const uint64_t* __restrict__ pool,
                                                    size_t rec_words, int W, const KnapEval* __restrict__ evals,
                                                    const uint64_t* __restrict__ inc_key, unsigned* __restrict__ masks,
                                                    int* __restrict__ counts) {
  __shared__ int s_cnt[kEvT / 32];
  const int nb = ctl->ev_nb;
  if (nb <= 0 || ctl->stop) return;
  const long long first = ctl->ev_first;
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (node < nb && evals[node].type == 2) {
    const double bv = evals[node].val;
    if (!ctl->has_inc || bv > ctl->inc_val) {
      keep = true;
    } else if (bv == ctl->inc_val) {
      const uint64_t* rec = pool + (size_t)(first + node) * rec_words;
      keep = knap_key_order(rec + 2 * (size_t)W, (int)rec[3 * (size_t)W], inc_key, ctl->inc_bits) <= 0;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    masks[blockIdx.x * (kEvT / 32) + warp] = m;
    s_cnt[warp] = __popc(m);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int q = 0; q < kEvT / 32; q++) c += s_cnt[q];
    counts[blockIdx.x] = c;
  }
}

// exclusive scan of the CTA counts (in place: counts[b] becomes the index of CTA b's first survivor, counts[ncta] the
// total), then stack bookkeeping, node budget and the next batch window
constexpr int kScanT = 1024;
__global__ void __launch_bounds__(kScanT) k_knap_scan(KnapCtl* ctl, int* counts) {
  __shared__ int s_part[kScanT];
  const int tid = threadIdx.x;
  const int nb = ctl->ev_nb;
  if (nb <= 0 || ctl->stop) {
    if (tid == 0) {
      ctl->ex_nj = 0;
      ctl->ev_nb = 0;
      ctl->stop = 1;
    }
    return;
  }
  const int ncta = (nb + kEvT - 1) / kEvT;
  const int per = (ncta + kScanT - 1) / kScanT;
  const int q0 = tid * per, q1 = min(ncta, q0 + per);
  int sum = 0;
  for (int q = q0; q < q1; q++) sum += counts[q];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < kScanT; o <<= 1) {  // inclusive scan of the per-thread sums
    const int add = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  int off = s_part[tid] - sum;
  for (int q = q0; q < q1; q++) {
    const int c = counts[q];
    counts[q] = off;
    off += c;
  }
  const int nj = s_part[kScanT - 1];
  if (tid == 0) {
    counts[ncta] = nj;
    const long long first = ctl->ev_first;
    if (first + 2LL * nj > ctl->pool_cap) {  // the batch cannot be expanded: report, keep the nodes where they are
      ctl->error = 1;
      ctl->stop = 1;
      ctl->ex_nj = 0;
      ctl->ev_nb = 0;
    } else {
      ctl->processed += nb;
      ctl->levels++;
      const long long open = first + 2LL * nj;
      ctl->open = open;
      ctl->ex_first = first;
      ctl->ex_nj = nj;
      ctl->ex_ncta = ncta;
      long long next = open < (long long)ctl->batch ? open : (long long)ctl->batch;
      if (ctl->max_nodes >= 0) {
        const long long left = ctl->max_nodes - ctl->processed;
        if (left < next) next = left < 0 ? 0 : left;
      }
      ctl->ev_nb = (int)next;
      ctl->ev_first = open - next;
      if (next == 0) ctl->stop = 1;
    }
  }
}

// surviving parents -> staging (their slots are about to be overwritten by children of other parents).  Job j is the
// j-th survivor in batch order: its evaluation CTA is found by binary search over the scanned counts, its lane from
// the ballot words.  parent[j] is written for k_knap_expand.  One WARP per job (records range from 80 bytes at
// n = 110 to 3.8 KB at n = 10^4: a CTA per job left 118 of 128 threads idle on the small ones).
__global__ void __launch_bounds__(128) k_knap_gather(const KnapCtl* __restrict__ ctl, const uint64_t* __restrict__ pool,
                                                     size_t rec_words, const int* __restrict__ counts,
                                                     const unsigned* __restrict__ masks,
                                                     long long* __restrict__ parent, uint64_t* __restrict__ stage) {
  const int nj = ctl->ex_nj;
  const long long first = ctl->ex_first;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int job = warp; job < nj; job += nwarps) {
    int node = -1;
    if (lane == 0) {
      int lo = 0, hi = ctl->ex_ncta;  // counts[lo] <= job < counts[hi]: counts is non-decreasing, counts[ncta] = nj
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (counts[mid] <= job) lo = mid; else hi = mid;
      }
      int r = job - counts[lo];  // rank inside CTA lo
      for (int q = 0; q < kEvT / 32; q++) {
        const unsigned m = masks[lo * (kEvT / 32) + q];
        const int c = __popc(m);
        if (r < c) {
          node = lo * kEvT + q * 32 + (__fns(m, 0, r + 1));
          break;
        }
        r -= c;
      }
      parent[job] = node;
    }
    node = __shfl_sync(0xffffffffu, node, 0);
    const uint64_t* src = pool + (size_t)(first + node) * rec_words;
    uint64_t* dst = stage + (size_t)job * rec_words;
    for (int t = lane; t < (int)rec_words; t += 32) dst[t] = src[t];
  }
}

// both children of one surviving node (its record at src): the x_k = 1 child below the x_k = 0 child (the zero child is
// DFS-first), each carrying the parent's fill state at its critical item k.  All 32 lanes of a warp.
__device__ __forceinline__ void knap_child_words(int t, uint64_t x, int W, int k, int depth, double cap, double base,
                                                 uint64_t& x0, uint64_t& x1) {
  x0 = x1 = x;
  const int word = t % W, sect = t / W;
  if (t < 3 * W) {
    if (sect == 0 && word == (k >> 6)) {  // fixmask: mark k fixed
      x0 |= 1ull << (k & 63);
      x1 |= 1ull << (k & 63);
    } else if (sect == 1 && word == (k >> 6)) {  // fixval
      x1 |= 1ull << (k & 63);
    } else if (sect == 2 && word == (depth >> 6)) {  // key: append the branch bit
      x1 |= 1ull << (depth & 63);
    }
  } else if (t == 3 * W) {
    x0 = x1 = (uint64_t)(depth + 1);
  } else if (t == 3 * W + 1) {
    x0 = knap_pack_state(k, 0);
    x1 = knap_pack_state(k, 1);
  } else if (t == 3 * W + 2) {
    x0 = x1 = (uint64_t)__double_as_longlong(cap);
  } else {
    x0 = x1 = (uint64_t)__double_as_longlong(base);
  }
}
__device__ __forceinline__ void knap_write_children(int lane, const uint64_t* __restrict__ src, uint64_t* one, uint64_t* zero,
                                                    int rec_words, int W, int k, double cap, double base) {
  const int depth = (int)src[3 * (size_t)W];
  for (int t = lane; t < rec_words; t += 32) {
    uint64_t x0, x1;
    knap_child_words(t, src[t], W, k, depth, cap, base, x0, x1);
    zero[t] = x0;
    one[t] = x1;
  }
}

// children of surviving nodes: job = index among the survivors (its parent record sits in stage[job]); two records
// per job.  One warp per job.
__global__ void __launch_bounds__(128) k_knap_expand(const KnapCtl* __restrict__ ctl, uint64_t* pool, size_t rec_words,
                                                     int W, const long long* __restrict__ parent,
                                                     const KnapEval* __restrict__ evals,
                                                     const uint64_t* __restrict__ stage) {
  const int nj = ctl->ex_nj;
  const long long dst_first = ctl->ex_first;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int job = warp; job < nj; job += nwarps) {
    const KnapEval ev = evals[parent[job]];
    uint64_t* one = pool + (size_t)(dst_first + 2 * (long long)job) * rec_words;
    knap_write_children(lane, stage + (size_t)job * rec_words, one, one + rec_words, (int)rec_words, W, ev.crit, ev.cap,
                        ev.base);
  }
}

// stage[j] = pool[offset + j * stride]: the records one rank keeps after every rank has expanded the same root
__global__ void k_knap_pick_stride(const uint64_t* __restrict__ pool, size_t rec_words, long long kept, int offset,
                                   int stride, uint64_t* __restrict__ stage) {
  for (long long j = blockIdx.x; j < kept; j += gridDim.x) {
    const uint64_t* src = pool + (size_t)(offset + j * stride) * rec_words;
    uint64_t* dst = stage + (size_t)j * rec_words;
    for (int t = threadIdx.x; t < (int)rec_words; t += blockDim.x) dst[t] = src[t];
  }
}

// start of a run: the host's view of the stack (it may have exported / imported records) and the node budget
__global__ void k_knap_start(KnapCtl* ctl, long long open, long long budget, int batch) {
  ctl->open = open;
  ctl->batch = batch;
  ctl->max_nodes = budget < 0 ? -1 : ctl->processed + budget;
  long long next = open < (long long)batch ? open : (long long)batch;
  if (budget >= 0 && budget < next) next = budget;
  ctl->ev_nb = (int)next;
  ctl->ev_first = open - next;
  ctl->ex_nj = 0;
  ctl->ex_first = open;
  ctl->stop = next == 0 ? 1 : 0;
  ctl->error = 0;
}

// incumbent handed in by the host (another GPU found it): its key is already in inc_key, its payload stays on the host
__global__ void k_knap_set_inc(KnapCtl* ctl, double val, int bits) {
  ctl->has_inc = 1;
  ctl->inc_val = val;
  ctl->inc_bits = bits;
  ctl->inc_crit = -2;
  ctl->inc_version++;
}


// ---- narrow levels: one thread-block cluster walks them without kernel boundaries ---------------------------------
// A level of the six-kernel pipeline costs ~45 us whatever its width (six dependent launches); the tree of cfg4
// (n = 10^4) has ~540 open nodes per level, so it ran at the speed of the launches.  While the WHOLE stack fits one
// thread per node of a cluster (16 CTAs x 256 threads, 8 x 256 where 16 is not available), this kernel loops over
// levels on its own: evaluation -> best candidate of every CTA into every CTA's shared memory (DSMEM) -> incumbent ->
// survivor ballots, CTA counts through DSMEM, prefix -> survivors to the staging area -> children, with a hardware
// cluster barrier where the pipeline has a kernel boundary.  Same stack discipline, same order, same incumbent
// rule as the pipeline, which takes over (same control block) as soon as a level outgrows the cluster, the node
// budget would cut a level, or max_levels is reached.
namespace cg = cooperative_groups;
constexpr int kNarrowT = 256;
constexpr int kNarrowMaxCtas = 16;
struct KnapCandN {
  double val;
  int idx;   // node index in the stack, -1 = none
  int crit;  // its critical item (-1 for a candidate that used every item)
};

__global__ void __launch_bounds__(kNarrowT, 1)
k_knap_narrow(KnapCtl* ctl, uint64_t* pool, size_t rec_words, int W, int n_items, const double* __restrict__ w,
              const double* __restrict__ v, uint64_t* stage, uint64_t* inc_key, uint64_t* inc_rec, int max_levels,
              unsigned long long* prof, int wv_in_smem) {
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int cap_nodes = ncta * kNarrowT;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kNarrowT / 32;
  __shared__ KnapCandN s_cand[kNarrowMaxCtas];  // written by every CTA of the cluster (slot = its rank)
  __shared__ int s_cnt[kNarrowMaxCtas];         // survivors per CTA, likewise
  __shared__ KnapCandN s_wc[NW];
  __shared__ int s_wcnt[NW];
  // the copy jobs (survivors) are dealt evenly to the CTAs, whatever CTA evaluated them: written through DSMEM by the
  // evaluating thread, slot = job - first job of the target CTA
  __shared__ int s_jnode[kNarrowT];     // stack position of the survivor
  __shared__ int s_jcrit[kNarrowT];     // its critical item and the fill state there (what the children inherit)
  __shared__ double s_jcap[kNarrowT], s_jbase[kNarrowT];
  __shared__ int s_depth[kNarrowT];     // its depth (read off the record while it is copied)

  // entry: the whole stack must be the next batch (uniform over the cluster: nobody has written the block yet)
  if (ctl->stop || ctl->error) return;
  long long open = ctl->open;
  if (open <= 0 || open > cap_nodes || ctl->ev_first != 0 || ctl->ev_nb != open) return;
  // weights and values on chip when they fit (2 x 80 KB at n = 10^4): every acquire of a cluster barrier empties L1, so
  // the walk's w[p] / v[p] would otherwise be an L2 round trip each, level after level
  extern __shared__ double s_wv[];
  const double* wq = w;
  const double* vq = v;
  if (wv_in_smem) {
    for (int t = tid; t < n_items; t += kNarrowT) {
      s_wv[t] = w[t];
      s_wv[n_items + t] = v[t];
    }
    wq = s_wv;
    vq = s_wv + n_items;
    __syncthreads();
  }
  long long processed = ctl->processed;
  const long long max_nodes = ctl->max_nodes, pool_cap = ctl->pool_cap;
  const int batch = ctl->batch;
  int has_inc = ctl->has_inc, inc_bits = ctl->inc_bits, inc_crit = ctl->inc_crit, inc_version = ctl->inc_version;
  double inc_val = ctl->inc_val;
  int levels = 0, error = 0;
  cluster.sync();  // everybody has read the block before CTA 0 may leave and rewrite it

  auto ahead = [&](double va, int a, double vb, int b) { return knap_cand_ahead(pool, rec_words, W, 0, va, a, vb, b); };
  long long t_last = prof ? clock64() : 0;  // LPR_KNAP_PROFILE=1: clocks of CTA 0 per section
  auto stamp = [&](int slot) {
    if (prof && crank == 0 && tid == 0) {
      const long long t = clock64();
      prof[slot] += (unsigned long long)(t - t_last);
      t_last = t;
    }
  };

  while (true) {
    const int n = (int)open;
    // the stack is dealt to the CTAs in contiguous blocks of B nodes (stack order = CTA-major order), B as small as the
    // level allows: 540 nodes are 16 x 64 threads walking, not 3 x 256 on three SMs
    const int B = min(kNarrowT, max(32, (((n + ncta - 1) / ncta) + 31) & ~31));
    const int node = tid < B ? crank * B + tid : n;
    // ---- relaxation, best candidate of the CTA --------------------------------------------------------------------
    KnapEval ev;
    ev.type = 0;
    ev.val = 0.0;
    ev.crit = -1;
    const uint64_t* rec = pool + (size_t)min(node, n) * rec_words;  // (only dereferenced when node < n)
    if (node < n) ev = knap_eval_node(rec, W, n_items, wq, vq);
    stamp(0);
    double cval = ev.type == 1 ? ev.val : 0.0;
    int cidx = (node < n && ev.type == 1) ? node : -1;
    int ccrit = ev.crit;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, cval, o);
      const int oi = __shfl_xor_sync(0xffffffffu, cidx, o);
      const int oc = __shfl_xor_sync(0xffffffffu, ccrit, o);
      if (ahead(ov, oi, cval, cidx)) {
        cval = ov;
        cidx = oi;
        ccrit = oc;
      }
    }
    if (lane == 0) {
      s_wc[wid].val = cval;
      s_wc[wid].idx = cidx;
      s_wc[wid].crit = ccrit;
    }
    __syncthreads();
    if (wid == 0) {
      KnapCandN c;
      c.val = 0.0;
      c.idx = -1;
      c.crit = -1;
      if (lane < NW) c = s_wc[lane];
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, c.val, o);
        const int oi = __shfl_xor_sync(0xffffffffu, c.idx, o);
        const int oc = __shfl_xor_sync(0xffffffffu, c.crit, o);
        if (ahead(ov, oi, c.val, c.idx)) {
          c.val = ov;
          c.idx = oi;
          c.crit = oc;
        }
      }
      if (lane < ncta) cluster.map_shared_rank(s_cand, lane)[crank] = c;
    }
    stamp(1);
    cluster.sync();
    stamp(2);
    // ---- incumbent (k_knap_inc's rule), evaluated by every warp from the same slots -------------------------------
    {
      KnapCandN c;
      c.val = 0.0;
      c.idx = -1;
      c.crit = -1;
      if (lane < ncta) c = s_cand[lane];
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, c.val, o);
        const int oi = __shfl_xor_sync(0xffffffffu, c.idx, o);
        const int oc = __shfl_xor_sync(0xffffffffu, c.crit, o);
        if (ahead(ov, oi, c.val, c.idx)) {
          c.val = ov;
          c.idx = oi;
          c.crit = oc;
        }
      }
      bool upd = false;
      if (c.idx >= 0) {
        const uint64_t* rb = pool + (size_t)c.idx * rec_words;
        upd = !has_inc || c.val > inc_val ||
              (c.val == inc_val && knap_key_order(rb + 2 * (size_t)W, (int)rb[3 * (size_t)W], inc_key, inc_bits) < 0);
      }
      if (upd) {  // uniform over the cluster
        const uint64_t* src = pool + (size_t)c.idx * rec_words;
        const int bits = (int)src[3 * (size_t)W];
        cluster.sync();  // every warp of every CTA has compared against the OLD key
        if (crank == 0) {
          for (int t = tid; t < (int)rec_words; t += kNarrowT) inc_rec[t] = src[t];
          for (int t = tid; t < W; t += kNarrowT) inc_key[t] = src[2 * (size_t)W + t];
        }
        has_inc = 1;
        inc_val = c.val;
        inc_bits = bits;
        inc_crit = c.crit;
        inc_version++;
        cluster.sync();  // the new key is in place before the survivor test reads it
      }
    }
    // ---- survivors (k_knap_flag's rule), counts of all CTAs, prefix ----------------------------------------------------
    bool keep = false;
    if (node < n && ev.type == 2) {
      if (!has_inc || ev.val > inc_val) keep = true;
      else if (ev.val == inc_val)
        keep = knap_key_order(rec + 2 * (size_t)W, (int)rec[3 * (size_t)W], inc_key, inc_bits) <= 0;
    }
    stamp(3);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_wcnt[wid] = __popc(m);
    __syncthreads();
    int before = 0, mine = 0;  // survivors of this CTA before my warp / in all of it
#pragma unroll
    for (int q = 0; q < NW; q++) {
      const int c = s_wcnt[q];
      if (q < wid) before += c;
      mine += c;
    }
    if (tid < ncta) cluster.map_shared_rank(s_cnt, tid)[crank] = mine;
    cluster.sync();
    stamp(4);
    int base = 0, nj = 0;
    for (int q = 0; q < ncta; q++) {
      const int c = s_cnt[q];
      if (q < crank) base += c;
      nj += c;
    }
    if (2LL * nj > pool_cap) {  // cannot happen below cap_nodes, kept for symmetry with k_knap_scan
      error = 1;
      break;
    }
    // survivors -> copy jobs: job = rank of the survivor in stack order, per jobs to every CTA
    const int per = (nj + ncta - 1) / ncta;
    if (keep) {
      const int job = base + before + __popc(m & ((1u << lane) - 1u));
      const int target = job / per, slot = job - target * per;
      cluster.map_shared_rank(s_jnode, target)[slot] = node;
      cluster.map_shared_rank(s_jcrit, target)[slot] = ev.crit;
      cluster.map_shared_rank(s_jcap, target)[slot] = ev.cap;
      cluster.map_shared_rank(s_jbase, target)[slot] = ev.base;
    }
    cluster.sync();
    const int j0 = crank * per;                      // my first job
    const int mine_jobs = max(0, min(per, nj - j0));  // and how many
    // ---- surviving parents -> staging (their slots are about to be overwritten) ------------------------------------
    // The CTA's survivors are consecutive jobs, so its share of the staging area is ONE flat array of mine x rec_words
    // words: all threads walk it with eight independent loads in flight each (a warp per record, one load at a time,
    // made a level cost more than the six launches it replaces).
    constexpr int kU = 8;
    const int rw = (int)rec_words;
    const int total = mine_jobs * rw;
    uint64_t* my_stage = stage + (size_t)j0 * rec_words;
    for (int b0 = 0; b0 < total; b0 += kNarrowT * kU) {
      uint64_t x[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int idx = b0 + u * kNarrowT + tid;
        if (idx < total) {
          const int r = idx / rw, t = idx - r * rw;
          x[u] = pool[(size_t)s_jnode[r] * rec_words + t];
          if (t == 3 * W) s_depth[r] = (int)x[u];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int idx = b0 + u * kNarrowT + tid;
        if (idx < total) my_stage[idx] = x[u];
      }
    }
    stamp(5);
    cluster.sync();
    stamp(6);
    // ---- children --------------------------------------------------------------------------------------------------------
    for (int b0 = 0; b0 < total; b0 += kNarrowT * kU) {
      uint64_t x[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int idx = b0 + u * kNarrowT + tid;
        if (idx < total) x[u] = my_stage[idx];
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int idx = b0 + u * kNarrowT + tid;
        if (idx < total) {
          const int r = idx / rw, t = idx - r * rw;
          uint64_t x0, x1;
          knap_child_words(t, x[u], W, s_jcrit[r], s_depth[r], s_jcap[r], s_jbase[r], x0, x1);
          uint64_t* one = pool + (size_t)(2 * (long long)(j0 + r)) * rec_words;
          one[t] = x1;
          one[rec_words + t] = x0;
        }
      }
    }
    processed += n;
    levels++;
    open = 2LL * nj;
    stamp(7);
    cluster.sync();
    stamp(8);  // children visible to every CTA; shared arrays free for the next level
    if (open <= 0 || open > cap_nodes || levels >= max_levels) break;
    if (open > batch) break;
    if (max_nodes >= 0 && processed + open > max_nodes) break;  // the budget cuts the next level: the pipeline's job
  }

  if (crank == 0 && tid == 0) {  // the block as k_knap_scan leaves it
    ctl->has_inc = has_inc;
    ctl->inc_val = inc_val;
    ctl->inc_bits = inc_bits;
    ctl->inc_crit = inc_crit;
    ctl->inc_version = inc_version;
    ctl->ex_nj = 0;
    ctl->ex_ncta = 0;
    if (error) {
      ctl->error = 1;
      ctl->stop = 1;
      ctl->ev_nb = 0;
    } else {
      ctl->open = open;
      ctl->processed = processed;
      ctl->levels += levels;
      ctl->ex_first = open;
      long long next = open < (long long)batch ? open : (long long)batch;
      if (max_nodes >= 0) {
        const long long left = max_nodes - processed;
        if (left < next) next = left < 0 ? 0 : left;
      }
      ctl->ev_nb = (int)next;
      ctl->ev_first = open - next;
      if (next == 0) ctl->stop = 1;
    }
  }
}

// selection of a candidate node, in ORIGINAL item ids (warp per call)
__global__ void k_knap_selection(const uint64_t* rec, int W, int n, double capacity, const double* w, int crit,
                                 const int* rank, uint8_t* chosen) {
  const uint32_t* mask32 = reinterpret_cast<const uint32_t*>(rec);
  const uint32_t* val32 = reinterpret_cast<const uint32_t*>(rec + W);
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const bool fixed = (mask32[p >> 5] >> (p & 31)) & 1u;
    const bool one = (val32[p >> 5] >> (p & 31)) & 1u;
    bool take = fixed ? one : (crit < 0 || p < crit);
    chosen[rank[p]] = take ? 1 : 0;
  }
  (void)capacity;
  (void)w;
}

// ---- DP arbiter: KnapsackBranchBoundSolver.Solve(int, int[], int[]) ----------------------------------
__global__ void k_dp_item(const long long* __restrict__ prev, long long* __restrict__ next, uint32_t* __restrict__ take,
                          int cap, int wi, long long vi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool t = false;
  if (c <= cap) {
    long long best = prev[c];
    if (wi <= c) {
      long long alt = prev[c - wi] + vi;
      if (alt > best) {
        best = alt;
        t = true;
      }
    }
    next[c] = best;
  }
  const unsigned b = __ballot_sync(0xffffffffu, t);
  if ((threadIdx.x & 31) == 0 && c <= cap + 31) take[c >> 5] = b;
}
__global__ void k_dp_backtrack(const uint32_t* take, size_t words_per_item, int n, int cap, const int* w,
                               uint8_t* chosen) {
  if (threadIdx.x || blockIdx.x) return;
  int c = cap;
  for (int i = n - 1; i >= 0; i--) {
    const uint32_t word = take[(size_t)i * words_per_item + (c >> 5)];
    const bool t = (word >> (c & 31)) & 1u;
    chosen[i] = t ? 1 : 0;
    if (t) c -= w[i];
  }
}

}  // namespace lpr

using namespace lpr;

struct lpr_knap {
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr;
  int n = 0, W = 0;
  size_t rec_words = 0;
  double capacity = 0.0;
  std::vector<int> rank;  // ranked position -> original id
  double *d_w = nullptr, *d_v = nullptr;
  int* d_rank = nullptr;
  uint64_t* pool = nullptr;
  long long pool_cap = 0, open = 0;  // `open` mirrors ctl->open between runs (export / import work on it)
  KnapEval* d_eval = nullptr;
  KnapCand* d_cands = nullptr;   // best candidate of every evaluation CTA
  unsigned* d_masks = nullptr;   // survivor ballots, one word per warp of the batch
  int* d_counts = nullptr;       // survivors per evaluation CTA, scanned in place (+ the total)
  long long* d_parent = nullptr;
  uint64_t* stage = nullptr;  // surviving parents of a level (children are written over the batch's slots)
  int batch = 0;
  unsigned long long* d_nprof = nullptr;  // LPR_KNAP_PROFILE=1
  int narrow_ctas = -1;  // cluster size of k_knap_narrow on this device: -1 not probed yet, 0 unavailable / switched off
  KnapCtl* d_ctl = nullptr;
  KnapCtl* h_ctl = nullptr;       // pinned copy, refreshed at every host synchronisation of lpr_knap_run
  uint64_t* d_inc_key = nullptr;  // incumbent key (device copy, read by k_knap_plan)
  uint64_t* d_inc_rec = nullptr;  // incumbent node record (selection is extracted from it on demand)
  uint8_t* d_chosen = nullptr;
  // host mirror of the incumbent, valid when seen_version == h_ctl->inc_version
  bool has_inc = false;
  double inc_val = -INFINITY;
  std::vector<uint64_t> inc_key;
  int inc_key_bits = 0;
  std::vector<uint8_t> inc_chosen;
  int seen_version = 0;
  int64_t processed = 0;
  // LPR_KNAP_PROFILE=1: printed by lpr_knap_destroy
  double t_run = 0;
  int64_t n_syncs = 0, n_levels = 0, n_incumbents = 0;
};

static inline double knap_now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int knap_key_cmp(const uint64_t* a, int abits, const uint64_t* b, int bbits) {
  return knap_key_order(a, abits, b, bbits);
}

// bring the host mirror of the incumbent up to date with the device (the device finds incumbents on its own)
static int knap_refresh_incumbent(lpr_knap* h) {
  if (h->seen_version == h->h_ctl->inc_version) return LPR_OK;
  const KnapCtl& c = *h->h_ctl;
  h->inc_key.assign(h->W, 0);
  LPR_CUDA(cudaMemcpyAsync(h->inc_key.data(), h->d_inc_key, sizeof(uint64_t) * h->W, cudaMemcpyDeviceToHost, h->stream));
  k_knap_selection<<<1, 256, 0, h->stream>>>(h->d_inc_rec, h->W, h->n, h->capacity, h->d_w, c.inc_crit, h->d_rank, h->d_chosen);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaMemcpyAsync(h->inc_chosen.data(), h->d_chosen, h->n, cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  h->has_inc = c.has_inc != 0;
  h->inc_val = c.inc_val;
  h->inc_key_bits = c.inc_bits;
  h->n_incumbents += c.inc_version - h->seen_version;
  h->seen_version = c.inc_version;
  return LPR_OK;
}

// cluster size for k_knap_narrow: 16 CTAs where the device grants the non-portable size, else 8 (LPR_KNAP_NARROW=0: none)
static size_t knap_narrow_smem(int n_items) {  // weights + values in shared memory when they fit beside the static arrays
  const size_t bytes = 2 * (size_t)n_items * sizeof(double);
  return bytes <= 200u * 1024u ? bytes : 0;
}
static int knap_narrow_probe(size_t smem) {
  const char* e = getenv("LPR_KNAP_NARROW");
  if (e && atoi(e) == 0) return 0;
  if (smem && cudaFuncSetAttribute(k_knap_narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (int ncta : {kNarrowMaxCtas, 8}) {
    if (ncta > 8 && cudaFuncSetAttribute(k_knap_narrow, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(kNarrowT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_knap_narrow, &cfg) == cudaSuccess && n >= 1) return ncta;
    cudaGetLastError();
  }
  return 0;
}

static int knap_launch_narrow(lpr_knap* h, int max_levels) {
  if (h->narrow_ctas < 0) {
    h->narrow_ctas = knap_narrow_probe(knap_narrow_smem(h->n));
    if (h->narrow_ctas > 0 && getenv("LPR_KNAP_PROFILE")) {
      LPR_CUDA(cudaMalloc(&h->d_nprof, 16 * sizeof(unsigned long long)));
      LPR_CUDA(cudaMemset(h->d_nprof, 0, 16 * sizeof(unsigned long long)));
    }
  }
  if (h->narrow_ctas == 0) return LPR_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h->narrow_ctas);
  cfg.blockDim = dim3(kNarrowT);
  cfg.dynamicSmemBytes = knap_narrow_smem(h->n);
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = h->narrow_ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t ce = cudaLaunchKernelEx(&cfg, k_knap_narrow, h->d_ctl, h->pool, h->rec_words, h->W, h->n,
                                            (const double*)h->d_w, (const double*)h->d_v, h->stage, h->d_inc_key,
                                            h->d_inc_rec, max_levels, h->d_nprof, cfg.dynamicSmemBytes ? 1 : 0);
  if (ce != cudaSuccess) return fail(LPR_E_CUDA, "knapsack cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  count_launch();
  return LPR_OK;
}

extern "C" {

int lpr_knap_destroy(lpr_knap* h) {
  if (!h) return LPR_OK;
  cudaSetDevice(h->device);
  if (getenv("LPR_KNAP_PROFILE"))
    fprintf(stderr, "[lpr_knap] nodes=%lld levels=%lld host syncs=%lld incumbents=%lld | run %.4fs\n",
            (long long)h->processed, (long long)h->n_levels, (long long)h->n_syncs, (long long)h->n_incumbents, h->t_run);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->d_nprof) {
    unsigned long long hp[16];
    if (cudaMemcpy(hp, h->d_nprof, sizeof(hp), cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "[lpr_knap] k_knap_narrow clocks of CTA 0: eval %llu | CTA candidate %llu | barrier %llu | incumbent + "
              "survivor test %llu | counts + barrier %llu | gather %llu | barrier %llu | children %llu | barrier %llu\n",
              hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7], hp[8]);
    cudaFree(h->d_nprof);
  }
  cudaFree(h->d_w); cudaFree(h->d_v); cudaFree(h->d_rank); cudaFree(h->pool); cudaFree(h->d_eval);
  cudaFree(h->d_parent); cudaFree(h->stage); cudaFree(h->d_inc_key); cudaFree(h->d_inc_rec); cudaFree(h->d_chosen);
  cudaFree(h->d_cands); cudaFree(h->d_masks); cudaFree(h->d_counts);
  cudaFree(h->d_ctl);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LPR_OK;
}

int lpr_knap_create(int device, double capacity, int n, const double* weights, const double* values,
                    lpr_knap** out) {
  if (!out) return fail(LPR_E_BADARG, "out is null");
  *out = nullptr;
  if (n < 1 || !weights || !values) return fail(LPR_E_BADARG, "bad knapsack instance (n=%d)", n);
  for (int i = 0; i < n; i++)
    if (!(weights[i] > 0.0)) return fail(LPR_E_BADARG, "weight %d must be positive", i);
  int rc = select_device(device);
  if (rc) return rc;
  lpr_knap* h = new (std::nothrow) lpr_knap();
  if (!h) return fail(LPR_E_NOMEM, "host allocation failed");
  h->device = device;
  h->sms = sm_count(device);
  h->n = n;
  h->W = (n + 63) / 64;
  h->rec_words = 3 * (size_t)h->W + kRecTail;
  h->capacity = capacity;
  // rank by value/weight descending, ties by lower original id (host side setup, O(n log n))
  h->rank.resize(n);
  std::iota(h->rank.begin(), h->rank.end(), 0);
  std::vector<double> ratio(n);
  for (int i = 0; i < n; i++) ratio[i] = values[i] / weights[i];
  std::stable_sort(h->rank.begin(), h->rank.end(), [&](int a, int b) { return ratio[a] > ratio[b]; });
  std::vector<double> rw(n), rv(n);
  for (int p = 0; p < n; p++) {
    rw[p] = weights[h->rank[p]];
    rv[p] = values[h->rank[p]];
  }
  const char* be = getenv("LPR_KNAP_BATCH");
  h->batch = be ? std::max(32, atoi(be)) : 65536;  // wide trees: 630 M nodes/s at 65536 against 323 M at 16384
  const char* pe = getenv("LPR_KNAP_POOL_MB");
  const size_t pool_bytes = (size_t)(pe ? std::max(16, atoi(pe)) : 4096) << 20;
  h->pool_cap = std::max<long long>((long long)(pool_bytes / (h->rec_words * 8)), 4LL * h->batch);
  cudaError_t e;
#define TRY(x)                                                                                   \
  if ((e = (x)) != cudaSuccess) {                                                                \
    lpr_knap_destroy(h);                                                                         \
    return fail(e == cudaErrorMemoryAllocation ? LPR_E_NOMEM : LPR_E_CUDA, "%s failed: %s", #x, \
                cudaGetErrorString(e));                                                          \
  }
  TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TRY(cudaMalloc(&h->d_w, sizeof(double) * n));
  TRY(cudaMalloc(&h->d_v, sizeof(double) * n));
  TRY(cudaMalloc(&h->d_rank, sizeof(int) * n));
  TRY(cudaMalloc(&h->pool, sizeof(uint64_t) * h->rec_words * (size_t)h->pool_cap));
  TRY(cudaMalloc(&h->stage, sizeof(uint64_t) * h->rec_words * (size_t)h->batch));
  TRY(cudaMalloc(&h->d_eval, sizeof(KnapEval) * h->batch));
  TRY(cudaMalloc(&h->d_parent, sizeof(long long) * h->batch));
  {
    const int ncta = (h->batch + kEvT - 1) / kEvT;
    TRY(cudaMalloc(&h->d_cands, sizeof(KnapCand) * ncta));
    TRY(cudaMalloc(&h->d_masks, sizeof(unsigned) * ncta * (kEvT / 32)));
    TRY(cudaMalloc(&h->d_counts, sizeof(int) * (ncta + 1)));
  }
  TRY(cudaMalloc(&h->d_inc_key, sizeof(uint64_t) * h->W));
  TRY(cudaMalloc(&h->d_inc_rec, sizeof(uint64_t) * h->rec_words));
  TRY(cudaMalloc(&h->d_chosen, n));
  TRY(cudaMalloc(&h->d_ctl, sizeof(KnapCtl)));
  TRY(cudaMallocHost(&h->h_ctl, sizeof(KnapCtl)));
  TRY(cudaMemsetAsync(h->d_inc_key, 0, sizeof(uint64_t) * h->W, h->stream));
  TRY(cudaMemsetAsync(h->d_inc_rec, 0, sizeof(uint64_t) * h->rec_words, h->stream));
  TRY(cudaMemcpyAsync(h->d_w, rw.data(), sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  TRY(cudaMemcpyAsync(h->d_v, rv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  TRY(cudaMemcpyAsync(h->d_rank, h->rank.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
  memset(h->h_ctl, 0, sizeof(KnapCtl));
  h->h_ctl->pool_cap = h->pool_cap;
  h->h_ctl->max_nodes = -1;
  h->h_ctl->batch = h->batch;
  h->h_ctl->inc_crit = -1;
  h->h_ctl->inc_val = -INFINITY;
  TRY(cudaMemcpyAsync(h->d_ctl, h->h_ctl, sizeof(KnapCtl), cudaMemcpyHostToDevice, h->stream));
  // root node: nothing fixed, empty key, fill state = (no parent item, whole capacity, no value)
  {
    std::vector<uint64_t> root(h->rec_words, 0);
    root[3 * (size_t)h->W + 1] = knap_pack_state(-1, 0);
    memcpy(&root[3 * (size_t)h->W + 2], &capacity, sizeof(double));
    TRY(cudaMemcpyAsync(h->pool, root.data(), sizeof(uint64_t) * h->rec_words, cudaMemcpyHostToDevice, h->stream));
    TRY(cudaStreamSynchronize(h->stream));
  }
#undef TRY
  h->open = 1;
  h->inc_chosen.assign(n, 0);
  h->inc_key.assign(h->W, 0);
  *out = h;
  return LPR_OK;
}

int lpr_knap_open_count(lpr_knap* h, int64_t* n) {
  if (!h || !n) return fail(LPR_E_BADARG, "null argument");
  *n = h->open;
  return LPR_OK;
}

// Levels are enqueued several at a time; the host only reads the control block between such groups (4, 8, ... 32
// levels), so a level costs six back-to-back launches and no round trip.  A group enqueued after the device has
// stopped (pool empty, budget reached) falls through: every kernel returns on ctl->stop / empty counts.
int lpr_knap_run_timed(lpr_knap* h, int64_t max_nodes, double max_seconds, int64_t* processed, int* status) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  const double t0 = knap_now();
  const long long before = h->h_ctl->processed;
  const int g_eval = (h->batch + kEvT - 1) / kEvT;
  const int g_exp = std::max(1, std::min(h->batch, h->sms * 8));
  k_knap_start<<<1, 1, 0, h->stream>>>(h->d_ctl, h->open, (long long)max_nodes, h->batch);
  LPR_LAUNCH_CHECK();
  int group = 4;
  // narrow stacks are walked by one cluster without kernel boundaries (k_knap_narrow): it returns at once when the
  // stack is wider than the cluster, and the pipeline below is a no-op once the device has stopped
  const int narrow_levels = max_seconds > 0.0 ? 64 : 4096;
  while (true) {
    if ((rc = knap_launch_narrow(h, narrow_levels))) return rc;
    for (int l = 0; l < group; l++) {
      k_knap_eval<<<g_eval, kEvT, 0, h->stream>>>(h->d_ctl, h->pool, h->rec_words, h->W, h->n, h->d_w, h->d_v, h->d_eval,
                                                   h->d_cands);
      LPR_LAUNCH_CHECK();
      k_knap_inc<<<1, kIncT, 0, h->stream>>>(h->d_ctl, h->pool, h->rec_words, h->W, h->d_eval, h->d_cands, h->d_inc_key,
                                             h->d_inc_rec);
      LPR_LAUNCH_CHECK();
      k_knap_flag<<<g_eval, kEvT, 0, h->stream>>>(h->d_ctl, h->pool, h->rec_words, h->W, h->d_eval, h->d_inc_key,
                                                   h->d_masks, h->d_counts);
      LPR_LAUNCH_CHECK();
      k_knap_scan<<<1, kScanT, 0, h->stream>>>(h->d_ctl, h->d_counts);
      LPR_LAUNCH_CHECK();
      k_knap_gather<<<g_exp, 128, 0, h->stream>>>(h->d_ctl, h->pool, h->rec_words, h->d_counts, h->d_masks, h->d_parent,
                                                  h->stage);
      LPR_LAUNCH_CHECK();
      k_knap_expand<<<g_exp, 128, 0, h->stream>>>(h->d_ctl, h->pool, h->rec_words, h->W, h->d_parent, h->d_eval, h->stage);
      LPR_LAUNCH_CHECK();
    }
    LPR_CUDA(cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(KnapCtl), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    h->n_syncs++;
    if (h->h_ctl->error) {
      h->open = h->h_ctl->open;
      return fail(LPR_E_CAPACITY, "knapsack node pool full (%lld records); raise LPR_KNAP_POOL_MB", h->pool_cap);
    }
    if (h->h_ctl->stop) break;
    if (max_seconds > 0.0 && knap_now() - t0 >= max_seconds) break;
    // a time slice is only checked between groups: keep the groups short then (a 65536-node level takes ~100 us)
    group = std::min(max_seconds > 0.0 ? 8 : 32, group * 2);
  }
  h->open = h->h_ctl->open;
  const int64_t done = h->h_ctl->processed - before;
  h->processed = h->h_ctl->processed;
  h->n_levels = h->h_ctl->levels;
  h->t_run += knap_now() - t0;
  if (processed) *processed = done;
  if (status) *status = (h->open > 0 && max_nodes >= 0 && done >= max_nodes) ? LPR_NODE_LIMIT : LPR_OPTIMAL;
  return LPR_OK;
}

// keep the open nodes whose position in the stack is == offset (mod stride) and drop the others: after every rank has
// expanded the same root for the same number of nodes (the level loop is deterministic, so the pools are identical),
// each keeps its own share -- a start-up partition without any transfer (csrc/multi_gpu.cu)
int lpr_knap_keep_stride(lpr_knap* h, int offset, int stride) {
  if (!h || stride < 1 || offset < 0 || offset >= stride) return fail(LPR_E_BADARG, "bad keep_stride arguments");
  int rc = select_device(h->device);
  if (rc) return rc;
  const long long kept = h->open > offset ? (h->open - offset + stride - 1) / stride : 0;
  if (kept > h->batch) return fail(LPR_E_CAPACITY, "keep_stride: %lld records exceed the staging area", kept);
  if (kept > 0) {
    k_knap_pick_stride<<<(int)std::min<long long>(kept, 1024), 128, 0, h->stream>>>(h->pool, h->rec_words, kept, offset,
                                                                                   stride, h->stage);
    LPR_LAUNCH_CHECK();
    LPR_CUDA(cudaMemcpyAsync(h->pool, h->stage, sizeof(uint64_t) * h->rec_words * (size_t)kept, cudaMemcpyDeviceToDevice,
                             h->stream));
  }
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  h->open = kept;
  return LPR_OK;
}

int lpr_knap_run(lpr_knap* h, int64_t max_nodes, int64_t* processed, int* status) {
  return lpr_knap_run_timed(h, max_nodes, 0.0, processed, status);
}

int lpr_knap_get_incumbent(lpr_knap* h, double* best, uint8_t* chosen, uint64_t* key, int* key_bits) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = knap_refresh_incumbent(h))) return rc;
  if (best) *best = h->has_inc ? h->inc_val : -INFINITY;
  if (chosen) memcpy(chosen, h->inc_chosen.data(), h->n);
  if (key && h->has_inc) memcpy(key, h->inc_key.data(), sizeof(uint64_t) * h->W);
  if (key_bits) *key_bits = h->has_inc ? h->inc_key_bits : -1;
  return LPR_OK;
}

int lpr_knap_set_incumbent(lpr_knap* h, double best, const uint8_t* chosen, const uint64_t* key, int key_bits) {
  if (!h || !chosen || key_bits < 0 || (key_bits > 0 && !key)) return fail(LPR_E_BADARG, "bad incumbent");
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = knap_refresh_incumbent(h))) return rc;
  std::vector<uint64_t> k(h->W, 0);
  if (key) memcpy(k.data(), key, sizeof(uint64_t) * h->W);
  if (!h->has_inc || best > h->inc_val ||
      (best == h->inc_val && knap_key_cmp(k.data(), key_bits, h->inc_key.data(), h->inc_key_bits) < 0)) {
    h->has_inc = true;
    h->inc_val = best;
    h->inc_key = k;
    h->inc_key_bits = key_bits;
    h->inc_chosen.assign(chosen, chosen + h->n);
    LPR_CUDA(cudaMemcpyAsync(h->d_inc_key, h->inc_key.data(), sizeof(uint64_t) * h->W, cudaMemcpyHostToDevice, h->stream));
    k_knap_set_inc<<<1, 1, 0, h->stream>>>(h->d_ctl, best, key_bits);
    LPR_LAUNCH_CHECK();
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    h->h_ctl->has_inc = 1;
    h->h_ctl->inc_val = best;
    h->h_ctl->inc_bits = key_bits;
    h->h_ctl->inc_crit = -2;
    h->h_ctl->inc_version++;
    h->seen_version = h->h_ctl->inc_version;
  }
  return LPR_OK;
}

// records are position independent: export = bottom (shallowest) records of the stack
int lpr_knap_export_nodes(lpr_knap* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes, int* n_exported) {
  if (!h || !buf || !bytes || !n_exported) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const size_t rb = sizeof(uint64_t) * h->rec_words;
  long long k = std::min<long long>(std::min<long long>(max_nodes, h->open), (long long)(buf_cap / (int64_t)rb));
  if (k < 0) k = 0;
  if (k > 0) {
    // on the handle's stream and waited for: a device-to-device cudaMemcpy returns before the copy has run, and the
    // caller hands `buf` to NCCL on another stream right away
    LPR_CUDA(cudaMemcpyAsync(buf, h->pool, rb * k, cudaMemcpyDefault, h->stream));  // buf: host or device
    // close the gap: move the top k records into the hole (order inside the pool does not matter,
    // the DFS keys decide ties)
    const long long rest = h->open - k;
    const long long mv = std::min(k, rest);
    if (mv > 0)
      LPR_CUDA(cudaMemcpyAsync(h->pool, h->pool + (size_t)(h->open - mv) * h->rec_words, rb * mv,
                               cudaMemcpyDeviceToDevice, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    h->open -= k;
  }
  *bytes = (int64_t)(rb * k);
  *n_exported = (int)k;
  return LPR_OK;
}

int lpr_knap_import_nodes(lpr_knap* h, const void* buf, int64_t bytes) {
  if (!h || (!buf && bytes > 0)) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const size_t rb = sizeof(uint64_t) * h->rec_words;
  if (bytes % (int64_t)rb) return fail(LPR_E_BADARG, "byte count is not a multiple of the record size");
  const long long k = bytes / (int64_t)rb;
  if (h->open + k > h->pool_cap) return fail(LPR_E_CAPACITY, "knapsack node pool full");
  if (k > 0) {
    LPR_CUDA(cudaMemcpyAsync(h->pool + (size_t)h->open * h->rec_words, buf, rb * k, cudaMemcpyDefault, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));  // `buf` may be reused by the caller as soon as this returns
  }
  h->open += k;
  return LPR_OK;
}

int lpr_knap_solve(int device, double capacity, int n, const double* weights, const double* values, int64_t max_nodes,
                   double* best, uint8_t* chosen, int64_t* nodes, int* status) {
  lpr_knap* h = nullptr;
  int rc = lpr_knap_create(device, capacity, n, weights, values, &h);
  if (rc) return rc;
  int64_t done = 0;
  int st = LPR_OPTIMAL;
  rc = lpr_knap_run(h, max_nodes, &done, &st);
  if (rc == LPR_OK) rc = knap_refresh_incumbent(h);
  if (rc == LPR_OK) {
    if (best) *best = h->has_inc ? h->inc_val : 0.0;
    if (chosen) memcpy(chosen, h->inc_chosen.data(), n);
    if (nodes) *nodes = done;
    if (status) *status = st;
  }
  lpr_knap_destroy(h);
  return rc;
}

int lpr_knap_dp(int device, int capacity, int n, const int* weights, const int* values, double* best,
                uint8_t* chosen) {
  if (n < 1 || !weights || !values || !best) return fail(LPR_E_BADARG, "bad DP arguments");
  int rc = select_device(device);
  if (rc) return rc;
  if (capacity < 0) capacity = 0;
  const size_t cells = (size_t)capacity + 1;
  const size_t wpi = (cells + 31) / 32 + 1;  // take-bit words per item
  long long *d_a = nullptr, *d_b = nullptr;
  uint32_t* d_take = nullptr;
  int* d_w = nullptr;
  uint8_t* d_ch = nullptr;
  cudaError_t e = cudaMalloc(&d_a, sizeof(long long) * cells);
  if (e == cudaSuccess) e = cudaMalloc(&d_b, sizeof(long long) * cells);
  if (e == cudaSuccess) e = cudaMalloc(&d_take, sizeof(uint32_t) * wpi * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&d_w, sizeof(int) * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_ch, n);
  if (e == cudaSuccess) e = cudaMemset(d_a, 0, sizeof(long long) * cells);
  if (e == cudaSuccess) e = cudaMemcpy(d_w, weights, sizeof(int) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int blocks = (int)((cells + 255) / 256);
    for (int i = 0; i < n; i++) {
      k_dp_item<<<blocks, 256>>>(d_a, d_b, d_take + (size_t)i * wpi, capacity, weights[i], (long long)values[i]);
      count_launch();
      std::swap(d_a, d_b);
    }
    k_dp_backtrack<<<1, 32>>>(d_take, wpi, n, capacity, d_w, d_ch);
    count_launch();
    long long bv = 0;
    e = cudaMemcpy(&bv, d_a + capacity, sizeof(long long), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && chosen) e = cudaMemcpy(chosen, d_ch, n, cudaMemcpyDeviceToHost);
    *best = (double)bv;
  }
  cudaFree(d_a); cudaFree(d_b); cudaFree(d_take); cudaFree(d_w); cudaFree(d_ch);
  if (e != cudaSuccess)
    return fail(e == cudaErrorMemoryAllocation ? LPR_E_NOMEM : LPR_E_CUDA, "knapsack DP failed: %s", cudaGetErrorString(e));
  return LPR_OK;
}

}  // extern "C"
