// K2: batched Riccati backward factorisation + forward substitution, one persistent CTA per QP instance.
//
// Replaces HpipmInterface::solve + getRiccatiFeedback/Feedforward/CostToGo
// (lib/ocs2_ros2/ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-455), i.e. HPIPM's unconstrained-QP path
// (one Riccati factorisation + one substitution; reg_prim on the Hessian diagonals).
//
// Data layout: see include/b200sqp.h section (1).  P and p live in shared memory for the whole backward sweep;
// each stage streams its (A|B), Q, S, R, q, r, b record from HBM exactly once in the backward sweep and (A|B), b once
// more in the forward sweep.
#pragma once
#include "dense.cuh"

namespace b200sqp {

struct QpDeviceView {
  int B, N, nx, numax;
  const double *A, *Bm, *b, *Q, *S, *R, *q, *r, *dx0;
  const int* nu;  // [B][N] or nullptr
  double *K, *kff, *P, *p, *dx, *du;
  int* status;  // [B]
  int keepP;
  double reg;
};

__host__ __device__ inline size_t riccati_smem_doubles(int nx, int numax) {
  const int nw = nx + numax;
  // P, Pn, AB, W, Y(S~), Rs, vectors: p, qv, rv, bv, v, dxv, duv
  return static_cast<size_t>(nx) * nx * 2 + static_cast<size_t>(nx) * nw * 2 + static_cast<size_t>(numax) * nx +
         static_cast<size_t>(numax) * numax + 5 * static_cast<size_t>(nx) + 2 * static_cast<size_t>(numax) + 8;
}

__global__ void __launch_bounds__(256, 1) riccati_kernel(QpDeviceView v) {
  extern __shared__ double sm[];
  const int inst = blockIdx.x;
  const int nx = v.nx, numax = v.numax, N = v.N;
  const int nw = nx + numax;
  double* P = sm;
  double* Pn = P + nx * nx;
  double* AB = Pn + nx * nx;     // nx x (nx+numax): [A | B]
  double* W = AB + nx * nw;      // P * [A | B]
  double* Y = W + nx * nw;       // numax x nx (ld numax): S~ then L^-1 S~
  double* Rs = Y + numax * nx;   // numax x numax
  double* pv = Rs + numax * numax;
  double* qv = pv + nx;
  double* bv = qv + nx;
  double* vv = bv + nx;
  double* xv = vv + nx;
  double* rv = xv + nx;          // numax
  double* uv = rv + numax;       // numax
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;

  const size_t iN = static_cast<size_t>(inst) * N, iN1 = static_cast<size_t>(inst) * (N + 1);
  // terminal stage: P_N = Q_N + reg I, p_N = q_N
  {
    const double* QN = v.Q + (iN1 + N) * nx * nx;
    for (int i = threadIdx.x; i < nx * nx; i += blockDim.x) P[i] = QN[i] + ((i % nx) == (i / nx) ? v.reg : 0.0);
    block_copy(nx, v.q + (iN1 + N) * nx, pv);
    __syncthreads();
    if (v.keepP) {
      block_copy(nx * nx, P, v.P + (iN1 + N) * nx * nx);
      block_copy(nx, pv, v.p + (iN1 + N) * nx);
    }
  }

  for (int k = N - 1; k >= 0; --k) {
    const int nu = v.nu ? v.nu[iN + k] : numax;
    const size_t sk = iN + k;
    // ---- load the stage record ---------------------------------------------------------------------------
    block_copy(nx * nx, v.A + sk * nx * nx, AB);
    if (nu > 0) block_copy(nx * nu, v.Bm + sk * nx * numax, AB + nx * nx);
    block_copy(nx * nx, v.Q + (iN1 + k) * nx * nx, Pn);
    if (nu > 0) {
      block_copy(numax * nx, v.S + sk * numax * nx, Y);
      block_copy(numax * numax, v.R + sk * numax * numax, Rs);
      block_copy(nu, v.r + sk * numax, rv);
    }
    block_copy(nx, v.q + (iN1 + k) * nx, qv);
    block_copy(nx, v.b + sk * nx, bv);
    __syncthreads();
    // ---- W = P [A|B],  v = P b + p ---------------------------------------------------------------------------
    block_gemm<4, 4, false, false>(nx, nx + nu, nx, 1.0, P, nx, AB, nx, W, nx);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = pv[i];
      for (int j = 0; j < nx; ++j) s = fma(P[i + j * nx], bv[j], s);
      vv[i] = s;
    }
    __syncthreads();
    // ---- Q~ = Q + A'W_A (+reg), S~ = S + B'W_A, R~ = R + B'W_B (+reg), q~ = q + A'v, r~ = r + B'v --------------
    block_gemm<4, 4, true, true>(nx, nx, nx, 1.0, AB, nx, W, nx, Pn, nx);
    if (nu > 0) {
      block_gemm<2, 4, true, true>(nu, nx, nx, 1.0, AB + nx * nx, nx, W, nx, Y, numax);
      block_gemm<2, 2, true, true>(nu, nu, nx, 1.0, AB + nx * nx, nx, W + nx * nx, nx, Rs, numax);
      block_gemv<true, true>(nu, nx, 1.0, AB + nx * nx, nx, vv, rv);
    }
    block_gemv<true, true>(nx, nx, 1.0, AB, nx, vv, qv);
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) Pn[i + i * nx] += v.reg;
    for (int i = threadIdx.x; i < nu; i += blockDim.x) Rs[i + i * numax] += v.reg;
    __syncthreads();
    if (nu > 0) {
      // ---- factorise R~ = L L', Y = L^-1 S~, y = L^-1 r~ ------------------------------------------------------
      warp_cholesky_lower(nu, Rs, numax, &ok);
      __syncthreads();
      block_trsm_lower(nu, nx, Rs, numax, Y, numax);
      if (threadIdx.x == blockDim.x - 1) {  // y = L^-1 r~ (single column)
        for (int i = 0; i < nu; ++i) {
          double s = rv[i];
          for (int j = 0; j < i; ++j) s = fma(-Rs[i + j * numax], rv[j], s);
          rv[i] = s / Rs[i + i * numax];
        }
      }
      __syncthreads();
      // ---- P = Q~ - Y'Y, p = q~ - Y'y ; K = -L^-T Y, k = -L^-T y ------------------------------------------------
      block_gemm<4, 4, true, true>(nx, nx, nu, -1.0, Y, numax, Y, numax, Pn, nx);
      block_gemv<true, true>(nx, nu, -1.0, Y, numax, rv, qv);
      block_trsm_lowerT_neg(nu, nx, Rs, numax, Y, numax, v.K + sk * numax * nx, numax);
      if (threadIdx.x == blockDim.x - 1) {
        double* kf = v.kff + sk * numax;
        for (int i = nu - 1; i >= 0; --i) {
          double s = rv[i];
          for (int j = i + 1; j < nu; ++j) s = fma(-Rs[j + i * numax], -kf[j], s);
          kf[i] = -(s / Rs[i + i * numax]);
        }
      }
    }
    __syncthreads();
    // symmetrise (the lower and upper triangles were accumulated in different orders) and rotate buffers
    for (int t = threadIdx.x; t < nx * nx; t += blockDim.x) {
      const int i = t % nx, j = t / nx;
      P[t] = 0.5 * (Pn[i + j * nx] + Pn[j + i * nx]);
    }
    block_copy(nx, qv, pv);
    __syncthreads();
    if (v.keepP) {
      block_copy(nx * nx, P, v.P + (iN1 + k) * nx * nx);
      block_copy(nx, pv, v.p + (iN1 + k) * nx);
    }
  }
  // ---- forward substitution -------------------------------------------------------------------------------------
  block_copy(nx, v.dx0 + static_cast<size_t>(inst) * nx, xv);
  __syncthreads();
  block_copy(nx, xv, v.dx + iN1 * nx);
  for (int k = 0; k < N; ++k) {
    const int nu = v.nu ? v.nu[iN + k] : numax;
    const size_t sk = iN + k;
    const double* Kk = v.K + sk * numax * nx;
    const double* Ak = v.A + sk * nx * nx;
    const double* Bk = v.Bm + sk * nx * numax;
    for (int i = threadIdx.x; i < numax; i += blockDim.x) {
      double s = 0.0;
      if (i < nu) {
        s = v.kff[sk * numax + i];
        for (int j = 0; j < nx; ++j) s = fma(Kk[i + j * numax], xv[j], s);
      }
      uv[i] = s;
      v.du[sk * numax + i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      double s = v.b[sk * nx + i];
      for (int j = 0; j < nx; ++j) s = fma(Ak[i + j * nx], xv[j], s);
      for (int j = 0; j < nu; ++j) s = fma(Bk[i + j * nx], uv[j], s);
      vv[i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      xv[i] = vv[i];
      v.dx[(iN1 + k + 1) * nx + i] = vv[i];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int bad = !ok;
    for (int i = 0; i < nx; ++i) bad |= !isfinite(xv[i]);
    v.status[inst] = bad;
  }
}

}  // namespace b200sqp
