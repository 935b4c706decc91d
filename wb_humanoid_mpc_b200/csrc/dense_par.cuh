// Dense fp64 helpers in "phase function" form: the caller passes its slice (tid, nt) of the block; no barriers inside.
#pragma once
#include "wb_model.cuh"

namespace b200sqp {

struct Par;  // {tid, nt} (wb_dynamics.cuh)

// C(MxN, ldc) = (ACC ? C : 0) + alpha * op(A) * B ; op(A) = A (MxK, lda) or A^T (A stored KxM, lda) ; B is KxN (ldb)
template <int TM, int TN, bool TRANS_A, bool ACC, class PAR>
HD void par_gemm(PAR P, int M, int N, int K, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                 double* __restrict__ C, int ldc) {
  const int tilesM = (M + TM - 1) / TM, tilesN = (N + TN - 1) / TN;
  for (int t = P.tid; t < tilesM * tilesN; t += P.nt) {
    const int i0 = (t % tilesM) * TM, j0 = (t / tilesM) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
    for (int k = 0; k < K; ++k) {
      double a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = (i0 + i < M) ? (TRANS_A ? A[k + (i0 + i) * lda] : A[(i0 + i) + k * lda]) : 0.0;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = (j0 + j < N) ? B[k + (j0 + j) * ldb] : 0.0;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (i0 + i < M && j0 + j < N) {
          double* c = &C[(i0 + i) + (j0 + j) * ldc];
          *c = ACC ? fma(alpha, acc[i][j], *c) : alpha * acc[i][j];
        }
  }
}

// y(M) = (ACC ? y : 0) + alpha * op(A) x
template <bool TRANS_A, bool ACC, class PAR>
HD void par_gemv(PAR P, int M, int K, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ x, double* __restrict__ y) {
  for (int i = P.tid; i < M; i += P.nt) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(TRANS_A ? A[k + i * lda] : A[i + k * lda], x[k], s);
    y[i] = ACC ? fma(alpha, s, y[i]) : alpha * s;
  }
}

}  // namespace b200sqp
