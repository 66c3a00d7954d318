// tableau.cuh -- device tableau handle shared by tableau.cu / cutting.cu / bb.cu.
#pragma once
#include "common.cuh"

namespace lpr {

// Device-resident solver state: written only by the single-CTA select kernels, read by sweeps.
struct TabState {
  int status;      // LPR_RUNNING / OPTIMAL / ...
  int enter;       // entering column of the pivot being applied (or -1)
  int leave;       // leaving row of the pivot being applied
  int next_enter;  // fused primal path: entering column of the NEXT pivot (-1 = optimal after this one)
  int do_sweep;    // 1 when the following sweep launch must apply (leave, enter)
  int cur;         // which colbuf holds the factor column of the pivot being applied
  int src;         // out-of-place (B&B) path: which of T/T2 holds the current tableau
  int phase;       // rule specific phase (SENS / BB: 0 dual, 1 primal)
  int have_prev;   // B&B: at least one pivot done (previous tableau is in the other buffer)
  int dropped;     // B&B: the last tableau was dropped (:392-400)
  long long npiv;
  long long max_piv;
  long long group_base;  // blocked path: npiv at the start of the current group of delayed pivots
  double pivot;
  double enter_val;  // pipelined path: T[0, enter] of the current tableau (f0 of the next pivot)
};

struct TabView {
  double* T;    // current tableau (in-place rules) or buffer 0 (out-of-place)
  double* T2;   // buffer 1 (out-of-place rules), may be null
  int ld;       // leading dimension in doubles (multiple of 16 => rows are 128 B aligned)
  int R, C;     // logical rows / cols (last col = RHS)
  double* col[2];  // factor column double buffer, each Rcap doubles
  double* rhs;     // RHS column copy (fused primal path)
  double* prow;    // normalised pivot row, ld doubles (padding = 0)
  int* basis;      // R-1 entries
  TabState* st;
  int* log;  // (row, col) pairs
  long long log_cap;
};

}  // namespace lpr

struct lpr_tab {
  int device = 0;
  cudaStream_t stream = nullptr;
  int R = 0, C = 0, Rcap = 0, Ccap = 0, ld = 0;
  double* T = nullptr;
  double* T2 = nullptr;  // lazily allocated second buffer (out-of-place B&B pivots)
  double* col[2] = {nullptr, nullptr};
  double* rhs = nullptr;
  double* prow = nullptr;
  int* basis = nullptr;
  lpr::TabState* st = nullptr;
  lpr::TabState* st_host = nullptr;  // pinned mirror
  // blocked (delayed-update) primal path: K pending rank-1 updates applied by one sweep
  int blk_k = 0;
  double* blk_pr = nullptr;    // K x ld   normalised pivot rows of the pending pivots
  double* blk_f = nullptr;     // K x Rcap pre-update factor columns of the pending pivots
  double* blk_row0 = nullptr;  // ld       mirror of the objective row (always current)
  double* blk_rhs[2] = {nullptr, nullptr};  // Rcap mirror of the RHS column, double buffered
  int* blk_p = nullptr;        // K        pivot rows of the pending pivots
  // pipelined delayed-update path (tableau_pipelined.cu): two group slots, two streams
  struct PipeRes {
    int k = 0;
    double* pr[2] = {nullptr, nullptr};   // K x ld   pivot rows of the group in this slot
    double* f[2] = {nullptr, nullptr};    // Rcap x K factor columns (K contiguous per row)
    int* pidx = nullptr;                  // 2 x K pivot rows
    int* count = nullptr;                 // 2 group sizes
    double* row0 = nullptr;               // ld  objective-row mirror
    double* rhs = nullptr;                // Rcap RHS mirror
    cudaStream_t s_sel = nullptr, s_sw = nullptr;
    int sw_sms = 0;                       // SMs the sweep stream may use (green-context partition), 0 = all
    cudaEvent_t ev_sel[2] = {nullptr, nullptr}, ev_sw[2] = {nullptr, nullptr}, ev_in = nullptr, ev_out = nullptr;
  } pipe;
  lpr::MinIdx* selcand = nullptr;    // per-CTA entering candidates of the multi-CTA select
  unsigned* ticket = nullptr;        // "last CTA done" counter
  int* log = nullptr;
  long long log_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evb[2] = {nullptr, nullptr};
  float last_ms = 0.f;
  float last_sweep_us = 0.f;  // average sweep-kernel duration of the last solve run with flag 8
  int sms = 148;
  lpr::TabView view() const {
    lpr::TabView v;
    v.T = T; v.T2 = T2; v.ld = ld; v.R = R; v.C = C;
    v.col[0] = col[0]; v.col[1] = col[1]; v.rhs = rhs; v.prow = prow;
    v.basis = basis; v.st = st; v.log = log; v.log_cap = log_cap;
    return v;
  }
};

namespace lpr {
// internal helpers implemented in tableau.cu
int tab_alloc(int device, int rows, int cols, int row_cap, int col_cap, lpr_tab** out);
int tab_ensure_log(lpr_tab* h, long long cap);
int tab_ensure_T2(lpr_tab* h);
void tab_pipe_free(lpr_tab* h);  // tableau_pipelined.cu
int tab_solve_internal(lpr_tab* h, int rule, int64_t max_pivots, int flags, int* status,
                       int64_t* n_pivots, int* pivot_log, int64_t log_cap);
}  // namespace lpr
