// K2 for the whole-body sizes (nx = 58, nu_max = 23), large batches: the Riccati recursion of HpipmInterface::solve
// (lib/ocs2_ros2/ocs2_sqp/hpipm_catkin/src/HpipmInterface.cpp:166-455) as TWO kernels.
//
//   riccati_bwd_kernel  backward factorisation, one persistent CTA per instance (216 KB of shared memory: one CTA per SM).
//                       Nine warps: warps 0-7 run the fp64 tensor-path contractions (DMMA m8n8k4), warp 8 is a helper warp that
//                       (a) streams the next stage record with cp.async.bulk + an mbarrier: one elected lane arms the barrier and issues SIX bulk
//                       copies (A, b, B, Q, q, S land densely, leading dimension 58 / 23, exactly as they lie in HBM; the 552 doubles of r and R,
//                       whose HBM addresses are only 8-byte aligned, ride as cp.async), (b) factorises R~ = L L' and inverts L, register resident, WHILE the
//                       GEMM warps form Q~ and S~ (the factorisation only needs R~, which is formed first and handed over through a named
//                       barrier), (c) writes the cost-to-go out when it is kept.
//                       Per stage:  W = P [A | b | B], W_b += p           (P is kept as its lower triangle: the operand loads pick (max, min))
//                                   R~ = R + B'W_B  ->  warp 8: L, L^-1   ||  Q~ = Q + A'W_A (lower tiles only), [S~ | r~] = [S | r] + B'[W_A | v]
//                                   Yl = L^-1 [S~ | r~]
//                                   [P | p] = [Q~ | q~] - Yl'Yl (lower tiles only), [K | k] = -L^-T Yl
//   riccati_fwd_kernel  forward substitution, one small CTA per instance (two resident per SM, so the whole batch is in flight and the
//                       second, partial wave of the backward kernel does not serialise it): du = K dx + k, dx+ = A dx + B du + b with the
//                       stage operands double-buffered in shared memory by cp.async.bulk + mbarrier.
//
// Same arithmetic as riccati_kernel<58, 23> up to the order of two additions (reg_prim is added to the diagonal of P after the
// update instead of before; symmetric entries are computed once instead of twice and averaged).
#pragma once
#include <cuda_pipeline.h>

#include "riccati.cuh"

namespace b200sqp {
namespace ricwb {

#ifdef B200SQP_PHASE_CLOCK
#define WB_TICK(slot, who)                                               \
  if (threadIdx.x == (who) && blockIdx.x == 0) {                         \
    const long long now_ = clock64();                                    \
    g_ricClk[slot][0] = __LINE__;                                        \
    g_ricClk[slot][1] += now_ - wbT_;                                    \
    wbT_ = now_;                                                         \
  }
#define WB_CLOCK_BEGIN()                                                 \
  long long wbT_ = clock64();                                            \
  if (threadIdx.x == 0 && blockIdx.x == 0)                               \
    for (int i_ = 0; i_ < 32; ++i_) g_ricClk[i_][0] = g_ricClk[i_][1] = 0;
#else
#define WB_TICK(slot, who)
#define WB_CLOCK_BEGIN()
#endif

constexpr int NXR = 58, NX1R = 59, NMR = 23;
constexpr int LWR = 60, LMR = 28;                 // leading dimensions of the operands the kernel produces itself: = 4 (mod 8), conflict-free DMMA fragment loads
constexpr int GEMM_WARPS = 8, BWD_THREADS = 32 * (GEMM_WARPS + 1);
constexpr int PQ_D = NXR * NX1R + 2;              // [P | p], dense (ld 58), lower triangle of P valid
constexpr int AB_D = NXR * (NX1R + NMR);          // [A | b | B], dense (ld 58): a 58 x 82 matrix
constexpr int W_D = LWR * (NX1R + NMR);           // W (ld 60); later Yl (ld 28) and [K | k] (ld 23)
constexpr int Y_D = NMR * NX1R + 1;               // [S | r], dense (ld 23)
constexpr int R_D = LMR * NMR;                    // R, L^-1 (ld 28)
constexpr int YL_D = LMR * NX1R;
constexpr int KOUT_OFF = YL_D;                    // [K | k] behind Yl inside the W region
constexpr int BWD_SMEM_D = 2 * PQ_D + 2 * AB_D + W_D + 2 * Y_D + 3 * R_D + 8;
constexpr size_t BWD_SMEM_BYTES = static_cast<size_t>(BWD_SMEM_D) * 8;

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (bytes: a multiple of 16; both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy accesses of this thread before, async-proxy accesses after (a buffer read with ld.shared is about to be refilled by a bulk copy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("barrier.cta.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- one DMMA work item: an 8-row strip x NG column tiles -------------------------------------------------------------------------------------
// C(tile tm, tiles tn[0..NG)) = (ACC ? C : 0) + alpha * op(A) * B on column-major operands; op(A) = A (M x K) or A' (A stored K x M).
// KC > 0: contraction length known at compile time (the k loop unrolls completely and the scheduler hoists the fragment loads).
// NG is a compile-time count (mma_item dispatches on the number of tiles) so that the unrolled DMMA stream carries no per-tile branches.
// SYM_A (only with !TRANS_A, M = K): A is symmetric and only its lower triangle is valid: element (m, k) is read from (max, min).
// vadd != nullptr: column vcol of the result additionally gets the vector vadd (W_b = P b + p without a separate pass).
// PERM_A / PERM_B: the operand has leading dimension 58 with the contraction index running fastest (A' stored K x M, or B): lane group fr then
// takes row / column perm8(fr) = (0, 2, 4, 6, 1, 3, 5, 7)[fr] of the tile instead of fr.  A half-warp's fragment load touches the doubles
// {k + 58 n : k = 0..3, n = four rows / columns}: with n = 0, 1, 2, 3 the bank offsets 10 n (mod 16) = 0, 10, 4, 14 make the 4-wide runs of
// n = 0 and n = 3 collide (a two-way conflict on every load); with n = 0, 2, 4, 6 (and 1, 3, 5, 7 in the other half-warp) they are 0, 4, 8, 12:
// conflict free.  Which tile row a lane group feeds is arbitrary as long as the result is stored accordingly, which the epilogue does.
__device__ __forceinline__ int perm8(int j) { return j < 4 ? 2 * j : 2 * (j - 4) + 1; }
template <bool TRANS_A, bool ACC, int NG, int KC, bool SYM_A, bool PERM_A, bool PERM_B>
__device__ __forceinline__ void mma_item_ng(int lane, int M, int N, int Krt, double alpha, const double* __restrict__ A, int lda, int tm,
                                            const double* __restrict__ B, int ldb, const int* tn, double* __restrict__ C, int ldc,
                                            const double* __restrict__ vadd, int vcol) {
  const int K = KC ? KC : Krt;
  const int fr = lane >> 2, fk = lane & 3;
  const int Kmain = K & ~3;
  const int ar = (tm << 3) + (PERM_A ? perm8(fr) : fr);
  const bool arok = ar < M;
  const int arc = arok ? ar : 0;
  const double* Ap = TRANS_A ? A + fk + arc * lda : A + arc + fk * lda;
  const double* ApU = A + fk + arc * lda;   // SYM_A: the mirrored element (k, m)
  const int astep = TRANS_A ? 4 : 4 * lda;
  const int frB = PERM_B ? perm8(fr) : fr;
  const int cj0 = PERM_B ? perm8(2 * fk) : 2 * fk, cj1 = PERM_B ? perm8(2 * fk + 1) : 2 * fk + 1;   // tile columns of this lane's two results
  const double* Bp[NG];
  bool bok[NG];
  double c0[NG], c1[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int bn = (tn[g] << 3) + frB;
    bok[g] = bn < N;
    Bp[g] = B + fk + (bok[g] ? bn : 0) * ldb;
    c0[g] = c1[g] = 0.0;
  }
#pragma unroll
  for (int k0 = 0; k0 < Kmain; k0 += 4) {
    double a;
    if (SYM_A) {
      const double* ps = (k0 + fk <= arc) ? Ap : ApU + k0;
      a = arok ? *ps : 0.0;
    } else {
      a = arok ? *Ap : 0.0;
    }
    Ap += astep;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const double b = bok[g] ? Bp[g][k0] : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[g]), "+d"(c1[g]) : "d"(a), "d"(b));
    }
  }
  if (Kmain < K) {
    const bool kok = Kmain + fk < K;
    double a = 0.0;
    if (arok && kok) {
      if (SYM_A) a = (Kmain + fk <= arc) ? *Ap : ApU[Kmain];
      else a = *Ap;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const double b = (bok[g] && kok) ? Bp[g][Kmain] : 0.0;
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[g]), "+d"(c1[g]) : "d"(a), "d"(b));
    }
  }
  if (arok) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int cn0 = (tn[g] << 3) + cj0, cn1 = (tn[g] << 3) + cj1;
      if (cn0 < N) {
        double* c = &C[ar + cn0 * ldc];
        double r = ACC ? fma(alpha, c0[g], *c) : alpha * c0[g];
        if (vadd && cn0 == vcol) r += vadd[ar];
        *c = r;
      }
      if (cn1 < N) {
        double* c = &C[ar + cn1 * ldc];
        double r = ACC ? fma(alpha, c1[g], *c) : alpha * c1[g];
        if (vadd && cn1 == vcol) r += vadd[ar];
        *c = r;
      }
    }
  }
}
template <bool TRANS_A, bool ACC, int NG, int KC, bool SYM_A = false, bool PERM_A = false, bool PERM_B = false>
__device__ __forceinline__ void mma_item(int lane, int M, int N, int Krt, double alpha, const double* __restrict__ A, int lda, int tm,
                                         const double* __restrict__ B, int ldb, const int (&tn)[NG], int cnt, double* __restrict__ C, int ldc,
                                         const double* __restrict__ vadd = nullptr, int vcol = -1) {
  // (warp-uniform dispatch; NG bounds the tile count)
  if (NG >= 4 && cnt == 4) mma_item_ng<TRANS_A, ACC, (NG >= 4 ? 4 : NG), KC, SYM_A, PERM_A, PERM_B>(lane, M, N, Krt, alpha, A, lda, tm, B, ldb, tn, C, ldc, vadd, vcol);
  else if (NG >= 3 && cnt == 3) mma_item_ng<TRANS_A, ACC, (NG >= 3 ? 3 : NG), KC, SYM_A, PERM_A, PERM_B>(lane, M, N, Krt, alpha, A, lda, tm, B, ldb, tn, C, ldc, vadd, vcol);
  else if (NG >= 2 && cnt == 2) mma_item_ng<TRANS_A, ACC, (NG >= 2 ? 2 : NG), KC, SYM_A, PERM_A, PERM_B>(lane, M, N, Krt, alpha, A, lda, tm, B, ldb, tn, C, ldc, vadd, vcol);
  else if (cnt == 1) mma_item_ng<TRANS_A, ACC, 1, KC, SYM_A, PERM_A, PERM_B>(lane, M, N, Krt, alpha, A, lda, tm, B, ldb, tn, C, ldc, vadd, vcol);
}

// Work items of a symmetric 58 x 59 update (the [P | p] shape): row strip m needs the column tiles 0..m of the lower triangle plus tile 7,
// which holds the vector column 58 (the entries of the upper triangle inside those tiles come for free and are never read).
// 13 items, 43 of 64 tiles; items 13..18 of the schedules below are the six 4-tile items of a 23 x 59 product (strip it/2, tile group it%2).
struct SymItem {
  unsigned char m, cnt, n[4];
};
__constant__ SymItem kSymItems[13] = {
    {7, 4, {0, 1, 2, 3}}, {7, 4, {4, 5, 6, 7}}, {6, 4, {0, 1, 2, 3}}, {6, 4, {4, 5, 6, 7}}, {5, 4, {0, 1, 2, 3}}, {4, 4, {0, 1, 2, 3}}, {3, 4, {0, 1, 2, 3}},
    {2, 4, {0, 1, 2, 7}}, {5, 3, {4, 5, 7, 0}},   {1, 3, {0, 1, 7, 0}},   {4, 2, {4, 7, 0, 0}},   {0, 2, {0, 7, 0, 0}},   {3, 1, {7, 0, 0, 0}}};
// P2 schedule (items per GEMM warp, -1 ends): warps 1-3 form R~ first (3 tiles each); warps 0 and 4 share their scheduler with the helper
// warp, which runs the latency-bound factorisation during this phase, and get little work.
__constant__ signed char kP2Sched[8][4] = {{7, 11, -1, -1}, {0, 13, -1, -1}, {1, 14, -1, -1}, {2, 15, -1, -1},
                                           {18, -1, -1, -1}, {3, 16, 8, -1}, {4, 17, 9, -1},  {5, 6, 10, 12}};
// the rest of W (item = 2 * row strip + {0: column tiles 0-3, 1: column tiles 4-6}) over the six worker warps
__constant__ signed char kP1bItems[6][3] = {{0, 1, 15}, {2, 3, -1}, {4, 5, -1}, {6, 7, 8}, {9, 10, 11}, {12, 13, 14}};
// P2 item order for the six worker warps (round robin): the fourteen 4-tile items first, the five short ones last
__constant__ signed char kP2Order[19] = {0, 1, 2, 3, 4, 5, 6, 7, 13, 14, 15, 16, 17, 18, 8, 9, 10, 11, 12};
// stages without inputs (event nodes): only the 13 symmetric items
__constant__ signed char kP2SchedNoInput[8][4] = {{0, 8, -1, -1}, {1, 9, -1, -1}, {2, 10, -1, -1}, {3, 11, -1, -1},
                                                  {4, 12, -1, -1}, {5, -1, -1, -1}, {6, -1, -1, -1}, {7, -1, -1, -1}};

template <int NMAX>
__device__ __forceinline__ void warp_chol_inverse_reg(int n, double* __restrict__ A, int lda, double reg, double* __restrict__ Linv, int ldl, int* ok);

__global__ void __launch_bounds__(BWD_THREADS, 1) riccati_bwd_kernel(QpDeviceView v) {
  extern __shared__ double sm[];
  const int inst = blockIdx.x;
  if (v.skip && v.skip[inst * v.skipStride]) return;
  const int N = v.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* PQ[2] = {sm, sm + PQ_D};
  double* AB[2] = {sm + 2 * PQ_D, sm + 2 * PQ_D + AB_D};
  double* W = AB[1] + AB_D;
  double* Y[2] = {W + W_D, W + W_D + Y_D};
  double* Rs[2] = {Y[1] + Y_D, Y[1] + Y_D + R_D};
  double* Linv = Rs[1] + R_D;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(Linv + R_D);   // [A | b | B], S of the next stage
  unsigned long long* barQ = bar + 1;                                            // [Q | q] of the next stage
  __shared__ int ok;
  const bool gemmWarp = warp < GEMM_WARPS;
  const size_t iN = static_cast<size_t>(inst) * N, iN1 = static_cast<size_t>(inst) * (N + 1);

  if (tid == 0) {
    ok = 1;
    mbar_init(bar, 1);
    mbar_init(barQ, 1);
    mbar_fence_init();
  }
  int cur = 0;
  // terminal stage: [P_N | p_N] = [Q_N + reg I | q_N]
  {
    const double* QN = v.Q + (iN1 + N) * NXR * NXR;
    for (int i = tid; i < NXR * NXR; i += blockDim.x) PQ[cur][i] = QN[i] + ((i % NXR == i / NXR) ? v.reg : 0.0);
    for (int i = tid; i < NXR; i += blockDim.x) PQ[cur][NXR * NXR + i] = v.q[(iN1 + N) * NXR + i];
  }
  __syncthreads();
  // the cost-to-go of stage kk, symmetric from its lower triangle
  auto storeP = [&](const double* Pc, int kk, int t0, int nt) {
    for (int i = t0; i < NXR * NXR; i += nt) {
      const int r = i % NXR, c = i / NXR;
      v.P[(iN1 + kk) * NXR * NXR + i] = r >= c ? Pc[r + c * NXR] : Pc[c + r * NXR];
    }
    for (int i = t0; i < NXR; i += nt) v.p[(iN1 + kk) * NXR + i] = Pc[NXR * NXR + i];
  };
  if (v.keepP) storeP(PQ[cur], N, tid, blockDim.x);
  // Stage record k -> operand set `set`, issued by the helper warp only, in three parts: the bulk copies of [A | b | B] and S (one lane) as soon
  // as that set is free; the bulk copies of [Q | q] once the buffer that takes them ([P | p] of the previous stage) is dead; the 8-byte cp.async
  // of r and R (whose HBM addresses are only 8-byte aligned) after the factorisation.
  auto prefetchBulk = [&](int k, int set) {
    const size_t sk = iN + k;
    const int nu = v.nu ? v.nu[sk] : NMR;
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, static_cast<unsigned>((NXR * NX1R + NXR * nu + (nu > 0 ? NMR * NXR : 0)) * 8));
      bulk_g2s(AB[set], v.A + sk * NXR * NXR, NXR * NXR * 8, bar);
      bulk_g2s(AB[set] + NXR * NXR, v.b + sk * NXR, NXR * 8, bar);
      if (nu > 0) {
        bulk_g2s(AB[set] + NXR * NX1R, v.Bm + sk * NXR * NMR, static_cast<unsigned>(NXR * nu * 8), bar);
        bulk_g2s(Y[set], v.S + sk * NMR * NXR, NMR * NXR * 8, bar);
      }
    }
  };
  auto prefetchQ = [&](int k, double* qdst) {
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(barQ, static_cast<unsigned>(NXR * NX1R * 8));
      bulk_g2s(qdst, v.Q + (iN1 + k) * NXR * NXR, NXR * NXR * 8, barQ);
      bulk_g2s(qdst + NXR * NXR, v.q + (iN1 + k) * NXR, NXR * 8, barQ);
    }
  };
  auto prefetchSmall = [&](int k, int set) {
    const size_t sk = iN + k;
    const int nu = v.nu ? v.nu[sk] : NMR;
    if (nu > 0) {
      if (lane < nu) __pipeline_memcpy_async(Y[set] + NMR * NXR + lane, v.r + sk * NMR + lane, 8);
      const double* Rg = v.R + sk * NMR * NMR;
      int c = 0, r = lane;   // (r, c) of element i = lane + 32 j, without divisions
      for (int i = lane; i < NMR * NMR; i += 32) {
        while (r >= NMR) {
          r -= NMR;
          ++c;
        }
        __pipeline_memcpy_async(Rs[set] + r + c * LMR, Rg + i, 8);
        r += 32;
      }
    }
    __pipeline_commit();
  };
  unsigned parity = 0, parityQ = 0;
  if (!gemmWarp) {
    prefetchBulk(N - 1, 0);
    prefetchQ(N - 1, PQ[1 - cur]);
    prefetchSmall(N - 1, 0);
  }
  WB_CLOCK_BEGIN()

  for (int k = N - 1; k >= 0; --k) {
    const int set = (N - 1 - k) & 1;
    const size_t sk = iN + k;
    const int nu = v.nu ? v.nu[sk] : NMR;
    double* Pc = PQ[cur];
    double* Pn = PQ[1 - cur];
    double* ABk = AB[set];
    double* Yk = Y[set];
    double* Rk = Rs[set];
    const double* Bk = ABk + NXR * NX1R;
    const int tilesN = (NX1R + nu + 7) >> 3;
    // ---- [A | b | B], S, r, R of this stage have landed ----------------------------------------------------------------------------------------
    if (!gemmWarp) __pipeline_wait_prior(0);
    mbar_wait(bar, parity);
    parity ^= 1;
    __syncthreads();
    WB_TICK(1, 0)
    WB_TICK(20, 256)
    // ---- P1a: the columns of W = P [A | b | B] that R~ needs first: column tiles 7.. = [A_56 A_57 | b | B] (column b additionally gets p) --------
    if (gemmWarp) {
      const int tn[4] = {7, 8, 9, 10};
      mma_item<false, false, 4, NXR, true, false, true>(lane, NXR, NX1R + nu, 0, 1.0, Pc, NXR, warp, ABk, NXR, tn, tilesN - 7, W, LWR, Pc + NXR * NXR, NXR);
    } else if (v.keepP && k < N - 1) {
      storeP(Pc, k + 1, lane, 32);   // (the cost-to-go of stage k+1 is complete and read-only until the [Q | q] copy below)
    }
    __syncthreads();
    WB_TICK(2, 0)
    WB_TICK(21, 256)
    // ---- P1b / P2: helper warp: stream [A | b | B], S of stage k-1, then factorise R~ as soon as three GEMM warps hand it over -- the
    //      factorisation (latency bound, one warp) runs WHILE the GEMM warps finish W (column tiles 0..6), and form Q~ (lower tiles) and [S~ | r~]
    if (!gemmWarp) {
      if (k > 0) prefetchBulk(k - 1, 1 - set);
      WB_TICK(22, 256)
      if (nu > 0) {
        named_sync(1, 128);
        WB_TICK(23, 256)
        warp_chol_inverse_reg<24>(nu, Rk, LMR, v.reg, Linv, LMR, &ok);
        WB_TICK(24, 256)
      }
      if (k > 0) prefetchSmall(k - 1, 1 - set);
      WB_TICK(27, 256)
    } else if ((warp & 3) != 0) {
      // worker warps 1,2,3,5,6,7: the warps 0 and 4 share their scheduler (and its fp64 / shuffle issue slots) with the helper warp and sit this
      // phase out while it factorises
      const int wk = warp - 1 - (warp > 4);   // 0..5
      if (nu > 0 && warp <= 3) {   // R~ = R + B'W_B
        const int tn[3] = {0, 1, 2};
        mma_item<true, true, 3, NXR, false, true, false>(lane, nu, nu, 0, 1.0, Bk, NXR, warp - 1, W + LWR * NX1R, LWR, tn, 3, Rk, LMR);
        __threadfence_block();
        named_arrive(1, 128);
      }
      // the rest of W: column tiles 0..6 of the eight row strips, 16 items over the six workers (the three that formed R~ take two each)
      for (int j = 0; j < 3; ++j) {
        const int it = kP1bItems[wk][j];
        if (it < 0) break;
        const int strip = it >> 1;
        if (it & 1) {
          const int tb[4] = {4, 5, 6, 0};
          mma_item<false, false, 4, NXR, true, false, true>(lane, NXR, NX1R + nu, 0, 1.0, Pc, NXR, strip, ABk, NXR, tb, 3, W, LWR);
        } else {
          const int ta[4] = {0, 1, 2, 3};
          mma_item<false, false, 4, NXR, true, false, true>(lane, NXR, NX1R + nu, 0, 1.0, Pc, NXR, strip, ABk, NXR, ta, 4, W, LWR);
        }
      }
      named_sync(2, 32 * 6);   // W is complete (barrier of the six workers)
      WB_TICK(7, 32)
      mbar_wait(barQ, parityQ);         // [Q | q] of this stage has landed in Pn
      const int nItems = nu > 0 ? 19 : 13;
      for (int it = wk; it < nItems; it += 6) {
        const int id = nu > 0 ? kP2Order[it] : it;
        if (id < 13) {
          const SymItem s = kSymItems[id];
          const int tn[4] = {s.n[0], s.n[1], s.n[2], s.n[3]};
          mma_item<true, true, 4, NXR, false, true, false>(lane, NXR, NX1R, 0, 1.0, ABk, NXR, s.m, W, LWR, tn, s.cnt, Pn, NXR);
        } else {
          const int tm = (id - 13) >> 1, g = ((id - 13) & 1) * 4;
          const int tn[4] = {g, g + 1, g + 2, g + 3};
          mma_item<true, true, 4, NXR, false, true, false>(lane, nu, NX1R, 0, 1.0, Bk, NXR, tm, W, LWR, tn, 4, Yk, NMR);
        }
      }
    }
    parityQ ^= 1;
    WB_TICK(8, 0)
    __syncthreads();
    WB_TICK(3, 0)
    WB_TICK(25, 256)
    // [P | p] of the previous stage is dead now: its buffer takes [Q | q] of stage k-1 (needed at P2 of the next stage)
    if (!gemmWarp && k > 0) prefetchQ(k - 1, Pc);
    double* Yl = W;
    double* Kout = W + KOUT_OFF;
    if (nu > 0) {
      // ---- P3: [Yl | yl] = L^-1 [S~ | r~] (into the W region) -----------------------------------------------------------------------------
      if (gemmWarp) {
        for (int it = warp; it < 6; it += GEMM_WARPS) {
          const int tm = it >> 1, g = (it & 1) * 4;
          const int tn[4] = {g, g + 1, g + 2, g + 3};
          mma_item<false, false, 4, 0>(lane, nu, NX1R, nu, 1.0, Linv, LMR, tm, Yk, NMR, tn, 4, Yl, LMR);
        }
      }
      __syncthreads();
      WB_TICK(4, 0)
      // ---- P4: [P | p] = [Q~ | q~] - Yl'[Yl | yl] (lower tiles) ; [K | k] = -L^-T [Yl | yl] -------------------------------------------------
      if (gemmWarp) {
        for (int it = warp; it < 19; it += GEMM_WARPS) {
          if (it < 8 || it >= 14) {
            const SymItem s = kSymItems[it < 8 ? it : it - 6];
            const int tn[4] = {s.n[0], s.n[1], s.n[2], s.n[3]};
            mma_item<true, true, 4, 0>(lane, NXR, NX1R, nu, -1.0, Yl, LMR, s.m, Yl, LMR, tn, s.cnt, Pn, NXR);
          } else {
            const int j = it - 8, tm = j >> 1, g = (j & 1) * 4;
            const int tn[4] = {g, g + 1, g + 2, g + 3};
            mma_item<true, false, 4, 0>(lane, nu, NX1R, nu, -1.0, Linv, LMR, tm, Yl, LMR, tn, 4, Kout, NMR);
          }
        }
      }
      __syncthreads();
      WB_TICK(5, 0)
    }
    // ---- P5: reg_prim on the diagonal, gains out -------------------------------------------------------------------------------------------
    if (tid < NXR) Pn[tid + tid * NXR] += v.reg;
    if (nu > 0) {
      for (int t = tid; t < NMR * NXR; t += blockDim.x) v.K[sk * NMR * NXR + t] = (t % NMR < nu) ? Kout[t] : 0.0;
      for (int t = tid; t < nu; t += blockDim.x) v.kff[sk * NMR + t] = Kout[NMR * NXR + t];
    }
    cur = 1 - cur;
    WB_TICK(6, 0)
    WB_TICK(26, 256)
    // (the barrier at the top of the next stage orders this phase before the next GEMM)
  }
  __syncthreads();
  if (v.keepP) storeP(PQ[cur], 0, tid, blockDim.x);
  if (tid == 0) v.status[inst] = !ok;
}

// warp_chol_inverse (dense.cuh) with reg added to the diagonal while loading
template <int NMAX>
__device__ __forceinline__ void warp_chol_inverse_reg(int n, double* __restrict__ A, int lda, double reg, double* __restrict__ Linv, int ldl, int* ok) {
  const int lane = threadIdx.x & 31;
  double a[NMAX], z[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    const int r = i > lane ? i : lane, c = i > lane ? lane : i;
    a[i] = (i < n && lane < n) ? A[r + c * lda] + (i == lane ? reg : 0.0) : ((i == lane) ? 1.0 : 0.0);
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
    const double ljk = a[k] * inv;
    z[k] *= inv;
#pragma unroll
    for (int i = k + 1; i < NMAX; ++i) {
      const double vv = __shfl_sync(0xffffffffu, a[i], k) * inv;
      a[i] = fma(-vv, ljk, a[i]);
      z[i] = fma(-vv, z[k], z[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NMAX; ++i)
    if (i < n && lane < n) Linv[i + lane * ldl] = z[i];
  if (lane == 0 && !good) *ok = 0;
}

// ---- forward substitution ---------------------------------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 128;
constexpr int FWD_A = NXR * NXR, FWD_B = NXR * NMR, FWD_K = NMR * NXR;
constexpr int FWD_SET = FWD_A + NXR + FWD_B + FWD_K + 24;   // A | b | B | K | k (padded to an even count)
constexpr size_t FWD_SMEM_BYTES = static_cast<size_t>(2 * FWD_SET) * 8;

__global__ void __launch_bounds__(FWD_THREADS) riccati_fwd_kernel(QpDeviceView v) {
  extern __shared__ double sm[];
  const int inst = blockIdx.x;
  if (v.skip && v.skip[inst * v.skipStride]) return;
  const int N = v.N, tid = threadIdx.x;
  __shared__ unsigned long long bar[2];
  __shared__ double xv[NXR], uv[24], part[4][64];
  const size_t iN = static_cast<size_t>(inst) * N, iN1 = static_cast<size_t>(inst) * (N + 1);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < NXR; i += blockDim.x) {
    const double x0 = v.dx0[static_cast<size_t>(inst) * NXR + i];
    xv[i] = x0;
    v.dx[iN1 * NXR + i] = x0;
  }
  __syncthreads();
  auto prefetch = [&](int k, int set) {   // one thread: four bulk copies (A, b, B, K are contiguous, 16-byte multiples); k rides as plain loads
    const size_t sk = iN + k;
    double* s = sm + set * FWD_SET;
    const int nu = v.nu ? v.nu[sk] : NMR;
    fence_proxy_async();
    mbar_expect_tx(&bar[set], static_cast<unsigned>((FWD_A + NXR + (nu > 0 ? FWD_B + FWD_K : 0)) * 8));
    bulk_g2s(s, v.A + sk * FWD_A, FWD_A * 8, &bar[set]);
    bulk_g2s(s + FWD_A, v.b + sk * NXR, NXR * 8, &bar[set]);
    if (nu > 0) {
      bulk_g2s(s + FWD_A + NXR, v.Bm + sk * FWD_B, FWD_B * 8, &bar[set]);
      bulk_g2s(s + FWD_A + NXR + FWD_B, v.K + sk * FWD_K, FWD_K * 8, &bar[set]);
    }
  };
  if (tid == 0) prefetch(0, 0);
  unsigned parity[2] = {0, 0};
  const int w = tid >> 5, lane = tid & 31;
  for (int k = 0; k < N; ++k) {
    const int set = k & 1;
    const size_t sk = iN + k;
    const int nu = v.nu ? v.nu[sk] : NMR;
    const double* s = sm + set * FWD_SET;
    const double* Ak = s;
    const double* bk = s + FWD_A;
    const double* Bk = s + FWD_A + NXR;
    const double* Kk = s + FWD_A + NXR + FWD_B;
    if (tid == 0 && k + 1 < N) prefetch(k + 1, 1 - set);   // (the other set was last read before the barrier that ended stage k-1)
    const double kffv = (w == 0 && lane < nu) ? v.kff[sk * NMR + lane] : 0.0;
    mbar_wait(&bar[set], parity[set]);
    parity[set] ^= 1;
    // du = K dx + k: row = lane (23 of 32), the 58 columns split over the four warps
    if (nu > 0) {
      double s0 = 0.0, s1 = 0.0;
      if (lane < nu) {
        const int c0 = w * 15, c1 = (c0 + 15 < NXR) ? c0 + 15 : NXR;
        int c = c0;
        for (; c + 1 < c1; c += 2) {
          s0 = fma(Kk[lane + c * NMR], xv[c], s0);
          s1 = fma(Kk[lane + (c + 1) * NMR], xv[c + 1], s1);
        }
        if (c < c1) s0 = fma(Kk[lane + c * NMR], xv[c], s0);
      }
      part[w][lane] = s0 + s1;
    }
    __syncthreads();
    if (tid < NMR) {
      const double r = (tid < nu) ? (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]) + kffv : 0.0;
      uv[tid] = r;
      v.du[sk * NMR + tid] = r;
    }
    __syncthreads();
    // dx+ = A dx + B du + b: rows 0..57 on 64 threads x 2 column halves of [A | B]
    {
      const int row = tid & 63, half = tid >> 6;
      double s0 = 0.0, s1 = 0.0;
      if (row < NXR) {
        if (half == 0) {
          for (int c = 0; c < 40; c += 2) {
            s0 = fma(Ak[row + c * NXR], xv[c], s0);
            s1 = fma(Ak[row + (c + 1) * NXR], xv[c + 1], s1);
          }
        } else {
          for (int c = 40; c < NXR; c += 2) {
            s0 = fma(Ak[row + c * NXR], xv[c], s0);
            s1 = fma(Ak[row + (c + 1) * NXR], xv[c + 1], s1);
          }
          for (int c = 0; c < nu; ++c) s0 = fma(Bk[row + c * NXR], uv[c], s0);
          s0 += bk[row];
        }
      }
      part[half][row] = s0 + s1;
    }
    __syncthreads();
    if (tid < NXR) {
      const double r = part[0][tid] + part[1][tid];
      xv[tid] = r;
      v.dx[(iN1 + k + 1) * NXR + tid] = r;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int bad = 0;
    for (int i = 0; i < NXR; ++i) bad |= !isfinite(xv[i]);
    if (bad) v.status[inst] = 1;
  }
}

}  // namespace ricwb
}  // namespace b200sqp
