// refactor.cuh -- workspace of the B^-1 refactorisation (refactor.cu), owned by the lpr_rev handle.
#pragma once
#include "common.cuh"

namespace lpr {

struct RefactorWs {
  int np = 0;  // padded order the buffers were sized for (0 = not allocated)
  double *Bm = nullptr, *Xp = nullptr, *Rm = nullptr;  // three np x np matrices
  double *W = nullptr;     // np x 64 scratch panel (pivot search)
  double *RK = nullptr;    // 64 x 2 np: D^-1 times the pivot rows of [M | X]
  double *Dinv = nullptr;  // 64 x 64
  double *d_res = nullptr;
  int *used = nullptr, *piv_rows = nullptr, *sw = nullptr, *singular = nullptr;
  unsigned* ticket = nullptr;
};
void refactor_ws_free(RefactorWs& ws);
int refactor_ws_ensure(RefactorWs& ws, int m);
int refactor_binv(cudaStream_t stream, int m, int n, const double* A, int ldA, double* Binv, int ldB, const int* basis,
                  RefactorWs& ws, int mode, double* residual_out, double* residual_after_out, double* flops_out,
                  int* path_out);

}  // namespace lpr
