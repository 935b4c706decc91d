// Dense fp64 helpers in "phase function" form: the caller passes its slice (tid, nt) of the block; no barriers inside.
#pragma once
#include "wb_model.cuh"

namespace b200sqp {

struct Par {  // the calling thread's slice of a phase
  int tid, nt;
};
// the same phase slice with the item assignment rotated by k threads: lets several item loops of one phase start on different threads
HD Par rot(Par P, int k) {
  int t = P.tid - (k % P.nt);
  if (t < 0) t += P.nt;
  return Par{t, P.nt};
}

// dst[0..n) = src[0..n) by the whole phase, four independent loads in flight per thread (global -> shared staging of a record)
HD void par_copy(Par P, double* __restrict__ dst, const double* __restrict__ src, int n) {
  int i = P.tid;
  for (; i + 3 * P.nt < n; i += 4 * P.nt) {
    const double a = src[i], b = src[i + P.nt], c = src[i + 2 * P.nt], d = src[i + 3 * P.nt];
    dst[i] = a;
    dst[i + P.nt] = b;
    dst[i + 2 * P.nt] = c;
    dst[i + 3 * P.nt] = d;
  }
  for (; i < n; i += P.nt) dst[i] = src[i];
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

namespace b200sqp {

// C(MxN, ldc) = (ACC ? C : 0) + alpha * op(A) * B ; op(A) = A (MxK, lda) or A^T (A stored KxM, lda) ; B is KxN (ldb), on the fp64 tensor path: DMMA (mma.sync.aligned.m8n8k4.f64) on column-major operands in shared or global memory.
// One work item = one 8-row strip x NG column tiles, so the A fragment is loaded once per k-step and reused NG times.
// Edges are handled by predicated (zero-filled) fragment loads and predicated stores; no padding requirements on the operands.
// On the host (development harness) the same contraction runs as scalar loops.
template <bool TRANS_A, bool ACC, int NG, class PAR>
HD void par_mma_gemm(PAR P, int M, int N, int K, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                     double* __restrict__ C, int ldc) {
#ifdef __CUDA_ARCH__
  const int warp = P.tid >> 5, nwarps = P.nt >> 5, lane = P.tid & 31;
  const int tilesM = (M + 7) >> 3, tilesN = (N + 7) >> 3;
  const int fr = lane >> 2, fk = lane & 3;
  const int Kmain = K & ~3;
  // work items (m-tile, group of NG n-tiles) are walked without integer division: warps stride over m-tiles inside each n-group
  int tm = warp, ng = 0;
  while (tm >= tilesM) {
    tm -= tilesM;
    ng += NG;
  }
  while (ng < tilesN) {
    const int ar = (tm << 3) + fr;
    const bool arok = ar < M;
    const double* Ap = TRANS_A ? A + fk + (arok ? ar : 0) * lda : A + (arok ? ar : 0) + fk * lda;
    const int astep = TRANS_A ? 4 : 4 * lda;
    const double* Bp[NG];
    bool bok[NG];
    double c0[NG], c1[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int bn = ((ng + g) << 3) + fr;
      bok[g] = bn < N;
      Bp[g] = B + fk + (bok[g] ? bn : 0) * ldb;
      c0[g] = c1[g] = 0.0;
    }
    for (int k0 = 0; k0 < Kmain; k0 += 4) {
      const double a = arok ? *Ap : 0.0;
      Ap += astep;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const double b = bok[g] ? Bp[g][k0] : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[g]), "+d"(c1[g])
                     : "d"(a), "d"(b));
      }
    }
    if (Kmain < K) {
      const bool kok = Kmain + fk < K;
      const double a = (arok && kok) ? *Ap : 0.0;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const double b = (bok[g] && kok) ? Bp[g][Kmain] : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[g]), "+d"(c1[g])
                     : "d"(a), "d"(b));
      }
    }
    if (arok) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int cn = ((ng + g) << 3) + 2 * fk;
        if (cn < N) {
          double* c = &C[ar + cn * ldc];
          *c = ACC ? fma(alpha, c0[g], *c) : alpha * c0[g];
        }
        if (cn + 1 < N) {
          double* c = &C[ar + (cn + 1) * ldc];
          *c = ACC ? fma(alpha, c1[g], *c) : alpha * c1[g];
        }
      }
    }
    tm += nwarps;
    while (tm >= tilesM) {
      tm -= tilesM;
      ng += NG;
    }
  }
#else
  for (int t = P.tid; t < M * N; t += P.nt) {
    const int i = t % M, j = t / M;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(TRANS_A ? A[k + i * lda] : A[i + k * lda], B[k + j * ldb], s);
    double* c = &C[i + j * ldc];
    *c = ACC ? fma(alpha, s, *c) : alpha * s;
  }
#endif
}

}  // namespace b200sqp
