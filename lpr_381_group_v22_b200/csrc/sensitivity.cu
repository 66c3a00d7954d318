// sensitivity.cu -- the tableau-side work of SensitivityAnalysis/SensitivityAnalyzer.cs kept on the device
// (SURVEY 8(f) row 1): RebuildBasicsFromTableau :706-723, the solution rebuild of ReOptimize :158-164,
// AddNewConstraintNonInteractive :609-659.  DualSimplexIfNeeded / ReOptimize themselves are
// lpr_tab_solve(rule = LPR_RULE_SENS) in tableau.cu.  The final tableau of a solve stays in HBM across
// "change -> resolve" instead of making a D2H / H2D round trip per menu action.
#include <algorithm>
#include <vector>

#include "select.cuh"
#include "tableau.cuh"

namespace lpr {

// GetBasicRow(col) :64-76 for every column: IsPivotColumn(i, col) holds iff row i is the ONLY constraint row
// with |T| > EPS, so the answer is that row when its entry is within EPS of 1, else -1 (two such rows exclude
// each other).  One warp per column.
__global__ void k_sens_basic_rows(TabView v, int* colrow) {
  const double EPS = 1e-9;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int R = v.R, C = v.C, ld = v.ld;
  for (int j = warp; j < C - 1; j += nw) {
    int big = 0, first = INT_MAX;  // rows with |T| > EPS: how many, the first
    for (int i = 1 + lane; i < R; i += 32) {
      const double t = TAT(v.T, ld, i, j);
      // a row "near 1" that is not > EPS cannot exist (EPS << 1); rows that are NaN count as neither
      if (fabs(t) > EPS) {
        big++;
        first = min(first, i);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      big += __shfl_xor_sync(0xffffffffu, big, o);
      first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    }
    int r = -1;
    if (big == 1 && fabs(TAT(v.T, ld, first, j) - 1.0) < EPS) r = first;
    // NaN entries: |NaN - 1| < EPS and |NaN| > EPS are both false, i.e. such rows never qualify and never
    // disqualify, exactly as in the reference's comparisons
    if (lane == 0) colrow[j] = r;
  }
}

// RebuildBasicsFromTableau :706-723: basis[i-1] = first column whose basic row is i
__global__ void k_sens_basis_init(int* basis, int m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) basis[i] = INT_MAX;
}
__global__ void k_sens_basis_scatter(const int* colrow, int ncol, int* basis) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncol; j += gridDim.x * blockDim.x) {
    const int r = colrow[j];
    if (r >= 1) atomicMin(basis + (r - 1), j);
  }
}
__global__ void k_sens_basis_fin(int* basis, int m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    if (basis[i] == INT_MAX) basis[i] = -1;
}

// ReOptimize :158-164
__global__ void k_sens_solution(TabView v, const int* colrow, double* x) {
  const int C = v.C, ld = v.ld;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C - 1; j += gridDim.x * blockDim.x) {
    const int r = colrow[j];
    x[j] = (r < 0) ? 0.0 : TAT(v.T, ld, r, C - 1);
  }
}

// AddNewConstraintNonInteractive :609-659 in place (needs one row and one column of headroom):
// new row R: coeff_j = -tech[j] + sum over basis positions (in order) of tech[basis[pos]] * T[pos+1, j], each
// product and sum rounded separately like the C# loop; the RHS column moves one to the right and the new
// slack column takes its place.
__global__ void k_sens_add_row(TabView v, const double* tech, double rhs_minus_ax) {
  const int R = v.R, C = v.C, ld = v.ld;
  double* T = v.T;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int j = gid; j < C - 1; j += nth) {
    double coeff = -tech[j];
    for (int pos = 0; pos < R - 1; pos++) {
      const int bc = v.basis[pos];
      coeff = __dadd_rn(coeff, __dmul_rn(tech[bc], TAT(T, ld, pos + 1, j)));
    }
    TAT(T, ld, R, j) = coeff;
  }
  for (int i = gid; i <= R; i += nth) {
    if (i < R) {
      TAT(T, ld, i, C) = TAT(T, ld, i, C - 1);
      TAT(T, ld, i, C - 1) = 0.0;  // new slack column (also newT[0, newSlackCol] = 0 :650)
    } else {
      TAT(T, ld, R, C - 1) = 1.0;
      TAT(T, ld, R, C) = rhs_minus_ax;
      for (int j = C + 1; j < ld; j++) TAT(T, ld, R, j) = 0.0;
      v.basis[R - 1] = C - 1;
    }
  }
}

static int sens_basic_rows(lpr_tab* h, int** colrow_out) {
  int* colrow = nullptr;
  LPR_CUDA(cudaMalloc(&colrow, sizeof(int) * std::max(1, h->C - 1)));
  const int blocks = std::max(1, std::min(h->sms * 8, (h->C - 1 + 7) / 8));
  k_sens_basic_rows<<<blocks, 256, 0, h->stream>>>(h->view(), colrow);
  count_launch();
  *colrow_out = colrow;
  return LPR_OK;
}

}  // namespace lpr

using namespace lpr;

extern "C" {

int lpr_tab_sens_rebuild_basis(lpr_tab* h) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (h->R < 2) return LPR_OK;
  int* colrow = nullptr;
  if ((rc = sens_basic_rows(h, &colrow))) return rc;
  const int m = h->R - 1;
  k_sens_basis_init<<<std::max(1, (m + 255) / 256), 256, 0, h->stream>>>(h->basis, m);
  k_sens_basis_scatter<<<std::max(1, (h->C - 1 + 255) / 256), 256, 0, h->stream>>>(colrow, h->C - 1, h->basis);
  k_sens_basis_fin<<<std::max(1, (m + 255) / 256), 256, 0, h->stream>>>(h->basis, m);
  count_launch(3);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(colrow);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "sens_rebuild_basis: %s", cudaGetErrorString(e));
  return LPR_OK;
}

int lpr_tab_sens_solution(lpr_tab* h, double* x) {
  if (!h || !x) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const int n = h->C - 1;
  int* colrow = nullptr;
  if ((rc = sens_basic_rows(h, &colrow))) return rc;
  double* dx = nullptr;
  cudaError_t e = cudaMalloc(&dx, sizeof(double) * n);
  if (e == cudaSuccess) {
    k_sens_solution<<<std::max(1, (n + 255) / 256), 256, 0, h->stream>>>(h->view(), colrow, dx);
    count_launch();
    e = cudaMemcpyAsync(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(dx);
  cudaFree(colrow);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "sens_solution: %s", cudaGetErrorString(e));
  return LPR_OK;
}

int lpr_tab_sens_add_constraint(lpr_tab* h, const double* tech, double rhs_minus_ax) {
  if (!h || !tech) return fail(LPR_E_BADARG, "null argument");
  if (h->R + 1 > h->Rcap || h->C + 1 > h->Ccap)
    return fail(LPR_E_CAPACITY, "no headroom for one more row and column (rows %d/%d, cols %d/%d)", h->R, h->Rcap,
                h->C, h->Ccap);
  int rc = select_device(h->device);
  if (rc) return rc;
  std::vector<int> basis(std::max(1, h->R - 1));
  if (h->R > 1) LPR_CUDA(cudaMemcpy(basis.data(), h->basis, sizeof(int) * (h->R - 1), cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->R - 1; i++)
    if (basis[i] < 0 || basis[i] >= h->C - 1)  // the reference indexes tech[basicVars[pos]] and would throw
      return fail(LPR_E_BADARG, "constraint row %d has no basic variable (RebuildBasicsFromTableau gave -1)", i + 1);
  double* dtech = nullptr;
  LPR_CUDA(cudaMalloc(&dtech, sizeof(double) * (h->C - 1)));
  cudaError_t e = cudaMemcpyAsync(dtech, tech, sizeof(double) * (h->C - 1), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    const int work = std::max(h->C - 1, h->R + 1);
    k_sens_add_row<<<std::max(1, std::min(h->sms * 8, (work + 127) / 128)), 128, 0, h->stream>>>(h->view(), dtech,
                                                                                                rhs_minus_ax);
    count_launch();
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(dtech);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "sens_add_constraint: %s", cudaGetErrorString(e));
  h->R += 1;
  h->C += 1;
  return LPR_OK;
}

}  // extern "C"
