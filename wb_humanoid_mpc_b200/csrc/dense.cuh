// Block-cooperative dense fp64 helpers on column-major operands (shared or global memory).
// Small fixed-size problems (<= 64 x 96): one CTA owns the operands, threads own register micro-tiles.
#pragma once
#include <cuda_runtime.h>

namespace b200sqp {

// C(MxN, ldc) = (ACC ? C : 0) + alpha * op(A) * B,  op(A) = A (MxK, lda) or A^T (A is KxM, lda) ; B is KxN (ldb)
// Threads of the CTA take TM x TN register tiles in a strided loop.  Caller synchronises.
template <int TM, int TN, bool TRANS_A, bool ACC>
__device__ __forceinline__ void block_gemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
                                           const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc) {
  const int tilesM = (M + TM - 1) / TM, tilesN = (N + TN - 1) / TN;
  for (int t = threadIdx.x; t < tilesM * tilesN; t += blockDim.x) {
    const int i0 = (t % tilesM) * TM, j0 = (t / tilesM) * TN;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
    const bool full = (i0 + TM <= M) && (j0 + TN <= N);
    if (full) {
      for (int k = 0; k < K; ++k) {
        double a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = TRANS_A ? A[k + (i0 + i) * lda] : A[(i0 + i) + k * lda];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = B[k + (j0 + j) * ldb];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    } else {
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

// y(M) = (ACC ? y : 0) + alpha * op(A) x ; one thread per output row (M <= ~64, K <= ~100)
template <bool TRANS_A, bool ACC>
__device__ __forceinline__ void block_gemv(int M, int K, double alpha, const double* __restrict__ A, int lda,
                                           const double* __restrict__ x, double* __restrict__ y) {
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(TRANS_A ? A[k + i * lda] : A[i + k * lda], x[k], s);
    y[i] = ACC ? fma(alpha, s, y[i]) : alpha * s;
  }
}

// In-place lower Cholesky of an n x n matrix (n <= 32) by warp 0; returns false in *ok on a non-positive pivot.
// Right-looking, lane i owns row i.
__device__ __forceinline__ void warp_cholesky_lower(int n, double* __restrict__ A, int lda, int* ok) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  bool good = true;
  for (int j = 0; j < n; ++j) {
    __syncwarp();
    double d = A[j + j * lda];
    if (!(d > 0.0)) {
      good = false;
      d = 1.0;
    }
    const double dj = sqrt(d);
    __syncwarp();
    double lij = 0.0;
    if (lane >= j && lane < n) {
      lij = (lane == j) ? dj : A[lane + j * lda] / dj;
      A[lane + j * lda] = lij;
    }
    __syncwarp();
    // trailing update: column c (> j), rows >= c : A[r][c] -= L[r][j] * L[c][j]
    if (lane > j && lane < n) {
      for (int c = j + 1; c <= lane; ++c) A[lane + c * lda] = fma(-lij, A[c + j * lda], A[lane + c * lda]);
    }
  }
  __syncwarp();
  if (lane == 0 && !good) *ok = 0;
}

// Fused Cholesky factorisation and triangular inverse of an n x n SPD matrix (n <= NMAX <= 32) by ONE warp, register resident and
// branch free.  Lane j owns the FULL symmetric column j of the matrix (a[]) and column j of L^-1 (z[]); rows / columns >= n are padded
// with the identity, so the loops are compile-time.  Step k: every lane forms its own L[j][k] = a[k] / sqrt(pivot) from its own register
// (symmetry), column k of L is broadcast once (one shuffle per row) and used twice, for the symmetric trailing update
// a[i] -= L[i][k] L[j][k] and for the right-looking inverse update z[i] -= L[i][k] z[k].  No shared-memory traffic, no division.
// Writes Linv (lower triangular, zeros above the diagonal); *ok = 0 on a non-positive pivot.
template <int NMAX>
__device__ __forceinline__ void warp_chol_inverse(int n, const double* __restrict__ A, int lda, double* __restrict__ Linv, int ldl, int* ok) {
  const int lane = threadIdx.x & 31;
  double a[NMAX], z[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    const int r = i > lane ? i : lane, c = i > lane ? lane : i;   // lower-triangle source of the symmetric entry (i, lane)
    a[i] = (i < n && lane < n) ? A[r + c * lda] : ((i == lane) ? 1.0 : 0.0);
    z[i] = (i == lane) ? 1.0 : 0.0;
  }
  bool good = true;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    double d = __shfl_sync(0xffffffffu, a[k], k);
    if (!(d > 0.0)) {
      good = false;
      d = 1.0;
    }
    const double inv = rsqrt(d);
    const double ljk = a[k] * inv;   // L[lane][k]
    z[k] *= inv;
#pragma unroll
    for (int i = k + 1; i < NMAX; ++i) {
      const double v = __shfl_sync(0xffffffffu, a[i], k) * inv;   // L[i][k]
      a[i] = fma(-v, ljk, a[i]);
      z[i] = fma(-v, z[k], z[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NMAX; ++i)
    if (i < n && lane < n) Linv[i + lane * ldl] = z[i];
  if (lane == 0 && !good) *ok = 0;
}

// X = L^-1 X for an n x m right-hand side (column-major, ldx); one thread per column
__device__ __forceinline__ void block_trsm_lower(int n, int m, const double* __restrict__ L, int ldl, double* __restrict__ X, int ldx) {
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    double* x = X + c * ldx;
    for (int i = 0; i < n; ++i) {
      double s = x[i];
      for (int k = 0; k < i; ++k) s = fma(-L[i + k * ldl], x[k], s);
      x[i] = s / L[i + i * ldl];
    }
  }
}
// Xout = -L^-T X ; one thread per column
__device__ __forceinline__ void block_trsm_lowerT_neg(int n, int m, const double* __restrict__ L, int ldl, const double* __restrict__ X,
                                                      int ldx, double* __restrict__ Xout, int ldo) {
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    const double* x = X + c * ldx;
    double* o = Xout + c * ldo;
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int k = i + 1; k < n; ++k) s = fma(-L[k + i * ldl], -o[k], s);  // o holds the negated solution
      o[i] = -(s / L[i + i * ldl]);
    }
  }
}

__device__ __forceinline__ void block_copy(int n, const double* __restrict__ src, double* __restrict__ dst) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace b200sqp
