// Warp- and block-level dense fp64 helpers used by the Riccati kernel (the CTA-wide GEMMs live in dense_par.cuh).
#pragma once
#include <cuda_runtime.h>

namespace b200sqp {

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

__device__ __forceinline__ void block_copy(int n, const double* __restrict__ src, double* __restrict__ dst) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace b200sqp
