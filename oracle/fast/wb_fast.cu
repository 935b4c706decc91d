// CPU BASELINE (test / benchmark infrastructure only -- never linked into libb200sqp.so, never reachable from the product path).
//
// The timed CPU arm of bench.py (`cpu_baseline`, `--impl reference`): one whole-body SQP iteration per instance on the host cores, with
// the cost structure of the reference's CPU path -- analytic (CppAD-like) Jacobians per shooting node on a pool of threads, one dense
// sequential Riccati sweep, value-only roll-outs for the filter line search (ocs2::SqpSolver::runImpl,
// lib/ocs2_ros2/ocs2_sqp/ocs2_sqp/src/SqpSolver.cpp:193-284).  The checker oracle (oracle/wb_problem.hpp) differentiates with dense
// 93-direction dual numbers, which is an order of magnitude slower than the reference's generated sparse code and made the CPU/GPU ratio
// meaningless; this file is the fast restatement used for TIMING: the node arithmetic is the same __host__ __device__ phase
// functions the CUDA kernels run (wb_node_a.inc / wb_node_b.inc / wb_rollout_body.inc, executed here by plain loops), the Riccati
// recursion is plain C.  tests/test_oracle_fast.py pins it on the checker oracle (same iterate, log and gains to 1e-9), so the timed arm
// and the checked arm are the same algorithm.
//
// Two modes: CPU-B "host throughput" = one instance per worker thread on all cores; CPU-A "reference-like
// latency" = one instance, its shooting nodes spread over `node_threads` workers (task.info nThreads 4), sequential Riccati.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../wb_humanoid_mpc_b200/csrc/wb_host.cuh"
#include "../../wb_humanoid_mpc_b200/csrc/wb_dynamics.cuh"
#include "../../wb_humanoid_mpc_b200/csrc/wb_lq.cuh"

using namespace b200sqp;

namespace {

constexpr double kWeakEpsF = 1e-9;  // numeric_traits::weakEpsilon
constexpr int NM = NUT_MAX;

struct NodeWs {  // per worker thread
  std::vector<double> a, b, r, mid;
  NodeWs() : a(lqWsDoubles()), b(pjWsDoubles()), r(roWsDoubles()), mid(Mid::SIZE) {}
};

struct Inst {  // one MPC instance: views into the batch arrays + its own QP storage
  int N = 0;
  const double *t, *x0, *swing, *impact, *arm, *xref;
  const uint8_t *event, *contact;
  double *x, *u;
  std::vector<double> A, Bt, b, Q, St, Rt, q, rt, Pu, Px, u0, perf, ls, dx, dut, du, K, kff, Kre, P, p;
  std::vector<int> nut;
  void alloc(int N_, bool gains) {
    N = N_;
    const size_t n = N, n1 = N + 1;
    A.assign(n * NX * NX, 0.0);
    Bt.assign(n * NX * NM, 0.0);
    b.assign(n * NX, 0.0);
    Q.assign(n1 * NX * NX, 0.0);
    St.assign(n * NM * NX, 0.0);
    Rt.assign(n * NM * NM, 0.0);
    q.assign(n1 * NX, 0.0);
    rt.assign(n * NM, 0.0);
    Pu.assign(n * NU * NM, 0.0);
    Px.assign(n * NU * NX, 0.0);
    u0.assign(n * NU, 0.0);
    perf.assign(n1 * 4, 0.0);
    ls.assign(n1 * 4, 0.0);
    dx.assign(n1 * NX, 0.0);
    dut.assign(n * NM, 0.0);
    du.assign(n * NU, 0.0);
    K.assign(n * NM * NX, 0.0);
    kff.assign(n * NM, 0.0);
    if (gains) Kre.assign(n * NU * NX, 0.0);
    nut.assign(n, 0);
  }
};

void loadNodeHost(const Inst& I, int k, NodeIn& n) {  // wb_solver.cuh loadNode
  n.xref = I.xref + static_cast<size_t>(k) * NX;
  n.event = I.event[k];
  n.terminal = (k == I.N);
  n.contact[0] = I.contact[2 * k];
  n.contact[1] = I.contact[2 * k + 1];
  for (int c = 0; c < 2; ++c) {
    for (int j = 0; j < 3; ++j) n.swing[c][j] = I.swing[(2 * k + c) * 3 + j];
    n.impact[c] = I.impact[2 * k + c];
  }
  n.armPhase = I.arm[k];
  n.dt = 0.0;
  if (k < I.N) {
    const double ts = I.t[k] + (I.event[k] == 2 ? kWeakEpsF : 0.0);
    const double te = I.t[k + 1] - (I.event[k + 1] == 1 ? kWeakEpsF : 0.0);
    n.dt = te - ts;
  }
}

// setupIntermediateNode / setupTerminalNode / setupEventNode + projectTranscription of node k (lq_dyn_kernel + lq_proj_kernel)
void lqNode(const WbDeviceModel& m, Inst& I, int k, NodeWs& ws) {
  const int N = I.N;
  NodeIn n;
  loadNodeHost(I, k, n);
  n.x = I.x + static_cast<size_t>(k) * NX;
  double* perf = I.perf.data() + static_cast<size_t>(k) * 4;
  if (k < N) {
    n.u = I.u + static_cast<size_t>(k) * NU;
    n.xnext = I.x + static_cast<size_t>(k + 1) * NX;
  }
  if (k == N) {
    double* Q = I.Q.data() + static_cast<size_t>(k) * NX * NX;
    double* q = I.q.data() + static_cast<size_t>(k) * NX;
    double c = 0.0;
    for (int i = 0; i < NX * NX; ++i) Q[i] = (i % NX == i / NX) ? m.Qfd[i % NX] : 0.0;
    for (int i = 0; i < NX; ++i) {
      const double dx = n.x[i] - n.xref[i];
      q[i] = m.Qfd[i] * dx;
      c += 0.5 * m.Qfd[i] * dx * dx;
    }
    perf[0] = c;
    perf[1] = perf[2] = perf[3] = 0.0;
    return;
  }
  const size_t st = k;
  NodeOut out{I.A.data() + st * NX * NX, I.Bt.data() + st * NX * NM, I.b.data() + st * NX, I.Q.data() + st * NX * NX, I.St.data() + st * NM * NX,
              I.Rt.data() + st * NM * NM, I.q.data() + st * NX, I.rt.data() + st * NM, I.Pu.data() + st * NU * NM, I.Px.data() + st * NU * NX,
              I.u0.data() + st * NU, I.nut.data() + st, perf, nullptr};
  if (n.event == 1) {
    double c = 0.0;
    for (int i = 0; i < NX * NX; ++i) {
      out.A[i] = (i % NX == i / NX) ? 1.0 : 0.0;
      out.Q[i] = 0.0;
    }
    for (int i = 0; i < NX; ++i) {
      const double df = n.x[i] - n.xnext[i];
      out.b[i] = df;
      out.q[i] = 0.0;
      c += df * df;
    }
    perf[0] = 0.0;
    perf[1] = c;
    perf[2] = perf[3] = 0.0;
    *out.nut = 0;
    return;
  }
  {
    LqWs s;
    lqWsMap(ws.a.data(), s);
    double* const mid = ws.mid.data();
    constexpr int NT = 128;
#define PHASE(...)                      \
  for (int t_ = 0; t_ < NT; ++t_) {     \
    Par P{t_, NT};                      \
    __VA_ARGS__                         \
  }
#include "../../wb_humanoid_mpc_b200/csrc/wb_node_a.inc"
#undef PHASE
  }
  {
    const double* const mid = ws.mid.data();
    const double dt = mid[Mid::META + 3];
    PjWs s;
    pjWsMap(ws.b.data(), s);
    constexpr int NT = 256;
#define PHASE(...)                      \
  for (int t_ = 0; t_ < NT; ++t_) {     \
    Par P{t_, NT};                      \
    __VA_ARGS__                         \
  }
#include "../../wb_humanoid_mpc_b200/csrc/wb_node_b.inc"
#undef PHASE
  }
}

// computePerformance of node k at (x + alpha dx, u + alpha du) (rollout_kernel)
void rolloutNode(const WbDeviceModel& m, Inst& I, int k, double alpha, NodeWs& ws) {
  const int N = I.N;
  RoWs r;
  roWsMap(ws.r.data(), r);
  NodeIn n;
  loadNodeHost(I, k, n);
  double* perfOut = I.ls.data() + static_cast<size_t>(k) * 4;
  for (int i = 0; i < NX; ++i) {
    r.xa[i] = std::fma(alpha, I.dx[static_cast<size_t>(k) * NX + i], I.x[static_cast<size_t>(k) * NX + i]);
    if (k < N) r.xna[i] = std::fma(alpha, I.dx[static_cast<size_t>(k + 1) * NX + i], I.x[static_cast<size_t>(k + 1) * NX + i]);
  }
  if (k < N)
    for (int i = 0; i < NU; ++i) r.ua[i] = std::fma(alpha, I.du[static_cast<size_t>(k) * NU + i], I.u[static_cast<size_t>(k) * NU + i]);
  n.x = r.xa;
  n.u = r.ua;
  n.xnext = r.xna;
  if (k == N || n.event == 1) {
    double c = 0.0;
    for (int i = 0; i < NX; ++i) {
      if (k == N) {
        const double dx = r.xa[i] - n.xref[i];
        c += 0.5 * m.Qfd[i] * dx * dx;
      } else {
        const double df = r.xa[i] - r.xna[i];
        c += df * df;
      }
    }
    perfOut[0] = (k == N) ? c : 0.0;
    perfOut[1] = (k == N) ? 0.0 : c;
    perfOut[2] = 0.0;
    return;
  }
  constexpr int NT = 128;
  DynWs& W = *r.dyn;
#define PHASE(...)                      \
  for (int t_ = 0; t_ < NT; ++t_) {     \
    Par P{t_, NT};                      \
    __VA_ARGS__                         \
  }
#include "../../wb_humanoid_mpc_b200/csrc/wb_rollout_body.inc"
#undef PHASE
}

// ---- dense helpers (column-major; the inner loops run down a column so that gcc vectorises them) -------------------------------------------
// C(m x n, ldc) += alpha * A(m x k, lda) * B(k x n, ldb)
inline void gemm_nn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  for (int j = 0; j < n; ++j) {
    double* c = C + static_cast<size_t>(j) * ldc;
    for (int l = 0; l < k; ++l) {
      const double s = alpha * B[l + static_cast<size_t>(j) * ldb];
      const double* a = A + static_cast<size_t>(l) * lda;
      for (int i = 0; i < m; ++i) c[i] += s * a[i];
    }
  }
}
// C(m x n, ldc) += alpha * A'(A is k x m, lda) * B(k x n, ldb)
inline void gemm_tn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  for (int j = 0; j < n; ++j) {
    const double* bcol = B + static_cast<size_t>(j) * ldb;
    for (int i = 0; i < m; ++i) {
      const double* a = A + static_cast<size_t>(i) * lda;
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += a[l] * bcol[l];
      C[i + static_cast<size_t>(j) * ldc] += alpha * s;
    }
  }
}

// HpipmInterface::solve for the unconstrained QP: Riccati backward factorisation + forward substitution (riccati.cuh / oracle/sqp.hpp).
// Returns false when a Cholesky pivot is not positive.
bool riccati(Inst& I, double reg, const double* dx0, bool keepP) {
  const int N = I.N, nx1 = NX + 1;
  std::vector<double> PQ(NX * nx1), Pn(NX * nx1), W(NX * (nx1 + NM)), AB(NX * (nx1 + NM)), Y(NM * nx1), R(NM * NM), L(NM * NM), Li(NM * NM), Yl(NM * nx1);
  if (keepP) {
    I.P.assign(static_cast<size_t>(N + 1) * NX * NX, 0.0);
    I.p.assign(static_cast<size_t>(N + 1) * NX, 0.0);
  }
  bool ok = true;
  // terminal: [P | p] = [Q_N + reg I | q_N]
  std::memcpy(PQ.data(), I.Q.data() + static_cast<size_t>(N) * NX * NX, sizeof(double) * NX * NX);
  for (int i = 0; i < NX; ++i) PQ[i + i * NX] += reg;
  std::memcpy(PQ.data() + NX * NX, I.q.data() + static_cast<size_t>(N) * NX, sizeof(double) * NX);
  auto storeP = [&](int k) {
    if (!keepP) return;
    std::memcpy(I.P.data() + static_cast<size_t>(k) * NX * NX, PQ.data(), sizeof(double) * NX * NX);
    std::memcpy(I.p.data() + static_cast<size_t>(k) * NX, PQ.data() + NX * NX, sizeof(double) * NX);
  };
  storeP(N);
  std::fill(I.K.begin(), I.K.end(), 0.0);
  std::fill(I.kff.begin(), I.kff.end(), 0.0);
  for (int k = N - 1; k >= 0; --k) {
    const size_t sk = k;
    const int nu = I.nut[k];
    const int nc = nx1 + nu;
    // [A | b | B]
    std::memcpy(AB.data(), I.A.data() + sk * NX * NX, sizeof(double) * NX * NX);
    std::memcpy(AB.data() + NX * NX, I.b.data() + sk * NX, sizeof(double) * NX);
    if (nu) std::memcpy(AB.data() + NX * nx1, I.Bt.data() + sk * NX * NM, sizeof(double) * NX * nu);
    // W = P [A | b | B]; W_b += p
    std::fill(W.begin(), W.begin() + static_cast<size_t>(NX) * nc, 0.0);
    gemm_nn(NX, nc, NX, 1.0, PQ.data(), NX, AB.data(), NX, W.data(), NX);
    for (int i = 0; i < NX; ++i) W[NX * NX + i] += PQ[NX * NX + i];
    // [Q~ | q~] = [Q | q] + A'[W_A | v]
    std::memcpy(Pn.data(), I.Q.data() + sk * NX * NX, sizeof(double) * NX * NX);
    std::memcpy(Pn.data() + NX * NX, I.q.data() + sk * NX, sizeof(double) * NX);
    gemm_tn(NX, nx1, NX, 1.0, AB.data(), NX, W.data(), NX, Pn.data(), NX);
    for (int i = 0; i < NX; ++i) Pn[i + i * NX] += reg;
    if (nu) {
      const double* Bk = AB.data() + NX * nx1;
      // [S~ | r~] = [S | r] + B'[W_A | v] ; R~ = R + B'W_B
      for (int c = 0; c < NX; ++c)
        for (int r_ = 0; r_ < nu; ++r_) Y[r_ + c * NM] = I.St[sk * NM * NX + r_ + c * NM];
      for (int r_ = 0; r_ < nu; ++r_) Y[r_ + NX * NM] = I.rt[sk * NM + r_];
      gemm_tn(nu, nx1, NX, 1.0, Bk, NX, W.data(), NX, Y.data(), NM);
      for (int c = 0; c < nu; ++c)
        for (int r_ = 0; r_ < nu; ++r_) R[r_ + c * NM] = I.Rt[sk * NM * NM + r_ + c * NM];
      gemm_tn(nu, nu, NX, 1.0, Bk, NX, W.data() + NX * nx1, NX, R.data(), NM);
      for (int i = 0; i < nu; ++i) R[i + i * NM] += reg;
      // R~ = L L'
      for (int j = 0; j < nu; ++j) {
        double d = R[j + j * NM];
        for (int l = 0; l < j; ++l) d -= L[j + l * NM] * L[j + l * NM];
        if (!(d > 0.0)) {
          ok = false;
          d = 1.0;
        }
        const double ljj = std::sqrt(d);
        L[j + j * NM] = ljj;
        for (int i = j + 1; i < nu; ++i) {
          double s = R[i + j * NM];
          for (int l = 0; l < j; ++l) s -= L[i + l * NM] * L[j + l * NM];
          L[i + j * NM] = s / ljj;
        }
      }
      // Yl = L^-1 [S~ | r~] (forward substitution, all columns)
      for (int c = 0; c < nx1; ++c) {
        for (int i = 0; i < nu; ++i) {
          double s = Y[i + c * NM];
          for (int l = 0; l < i; ++l) s -= L[i + l * NM] * Yl[l + c * NM];
          Yl[i + c * NM] = s / L[i + i * NM];
        }
      }
      // [P | p] = [Q~ | q~] - Yl'[Yl | yl] ; [K | k] = -L^-T [Yl | yl]
      gemm_tn(NX, nx1, nu, -1.0, Yl.data(), NM, Yl.data(), NM, Pn.data(), NX);
      for (int c = 0; c < nx1; ++c) {
        double kc[NM];
        for (int i = nu - 1; i >= 0; --i) {
          double s = Yl[i + c * NM];
          for (int l = i + 1; l < nu; ++l) s -= L[l + i * NM] * kc[l];
          kc[i] = s / L[i + i * NM];
        }
        if (c < NX)
          for (int i = 0; i < nu; ++i) I.K[sk * NM * NX + i + c * NM] = -kc[i];
        else
          for (int i = 0; i < nu; ++i) I.kff[sk * NM + i] = -kc[i];
      }
    }
    // symmetrise, rotate
    for (int j = 0; j < NX; ++j)
      for (int i = j + 1; i < NX; ++i) {
        const double mavg = 0.5 * (Pn[i + j * NX] + Pn[j + i * NX]);
        Pn[i + j * NX] = mavg;
        Pn[j + i * NX] = mavg;
      }
    PQ.swap(Pn);
    storeP(k);
  }
  // forward substitution
  std::memcpy(I.dx.data(), dx0, sizeof(double) * NX);
  for (int k = 0; k < N; ++k) {
    const size_t sk = k;
    const int nu = I.nut[k];
    const double* xk = I.dx.data() + sk * NX;
    double* ut = I.dut.data() + sk * NM;
    for (int i = 0; i < NM; ++i) ut[i] = 0.0;
    for (int i = 0; i < nu; ++i) ut[i] = I.kff[sk * NM + i];
    for (int c = 0; c < NX; ++c)
      for (int i = 0; i < nu; ++i) ut[i] += I.K[sk * NM * NX + i + c * NM] * xk[c];
    double* xn = I.dx.data() + (sk + 1) * NX;
    for (int i = 0; i < NX; ++i) xn[i] = I.b[sk * NX + i];
    for (int c = 0; c < NX; ++c) {
      const double s = xk[c];
      const double* a = I.A.data() + sk * NX * NX + static_cast<size_t>(c) * NX;
      for (int i = 0; i < NX; ++i) xn[i] += a[i] * s;
    }
    for (int c = 0; c < nu; ++c) {
      const double s = ut[c];
      const double* bcol = I.Bt.data() + sk * NX * NM + static_cast<size_t>(c) * NX;
      for (int i = 0; i < NX; ++i) xn[i] += bcol[i] * s;
    }
  }
  for (int i = 0; i < NX; ++i) ok = ok && std::isfinite(I.dx[static_cast<size_t>(N) * NX + i]);
  return ok;
}

// run fn(k, ws) for k in [0, count) on `threads` workers (threads <= 1: inline)
template <class F>
void parallelNodes(int count, int threads, std::vector<NodeWs>& ws, F fn) {
  if (threads <= 1) {
    for (int k = 0; k < count; ++k) fn(k, ws[0]);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      for (int k = next++; k < count; k = next++) fn(k, ws[t]);
    });
  for (auto& th : pool) th.join();
}

struct Timers {
  double lq = 0, qp = 0, ls = 0;
};
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// SqpSolver::runImpl for one instance; log: [maxIter][16] (b200sqp_iter_log layout); returns the number of iterations, -1 on QP failure
int solveOne(const WbDeviceModel& m, Inst& I, const b200sqp_settings& st, int nodeThreads, std::vector<NodeWs>& ws, double* log, Timers& tm) {
  const int N = I.N;
  double dx0[NX];
  int iter = 0;
  for (; iter < st.sqp_iteration;) {
    double t0 = now();
    parallelNodes(N + 1, nodeThreads, ws, [&](int k, NodeWs& w) { lqNode(m, I, k, w); });
    // baseline PerformanceIndex (prep_kernel mode 0)
    double baseCost = 0, baseDyn = 0, baseEq = 0;
    for (int k = 0; k <= N; ++k) {
      baseCost += I.perf[4 * k];
      baseDyn += I.perf[4 * k + 1];
      baseEq += I.perf[4 * k + 2];
    }
    for (int i = 0; i < NX; ++i) {
      dx0[i] = I.x0[i] - I.x[i];
      baseDyn = std::fma(dx0[i], dx0[i], baseDyn);
    }
    const double baseMerit = baseCost;
    double t1 = now();
    tm.lq += t1 - t0;
    if (!riccati(I, st.reg_prim, dx0, st.create_value_function != 0)) return -1;
    // remapProjectedInput / remapProjectedGain, Armijo metric, norms (remap_kernel)
    double arm = 0, dxn = 0, dun = 0;
    for (int k = 0; k < N; ++k) {
      const size_t sk = k;
      const int nut = I.nut[k];
      const double* dx = I.dx.data() + sk * NX;
      const double* dut = I.dut.data() + sk * NM;
      double* du = I.du.data() + sk * NU;
      for (int i = 0; i < NU; ++i) {
        double s = 0.0;
        if (nut > 0) {
          s = I.u0[sk * NU + i];
          for (int j = 0; j < nut; ++j) s = std::fma(I.Pu[sk * NU * NM + i + NU * j], dut[j], s);
          for (int j = 0; j < NX; ++j) s = std::fma(I.Px[sk * NU * NX + i + NU * j], dx[j], s);
        }
        du[i] = s;
        dun = std::fma(s, s, dun);
      }
      for (int i = 0; i < NX; ++i) {
        arm = std::fma(I.q[sk * NX + i], dx[i], arm);
        dxn = std::fma(dx[i], dx[i], dxn);
      }
      for (int i = 0; i < nut; ++i) arm = std::fma(I.rt[sk * NM + i], dut[i], arm);
      if (!I.Kre.empty()) {
        double* Ko = I.Kre.data() + sk * NU * NX;
        for (int it = 0; it < NU * NX; ++it) {
          double s = 0.0;
          if (nut > 0) {
            const int i = it % NU, j = it / NU;
            s = I.Px[sk * NU * NX + it];
            for (int l = 0; l < nut; ++l) s = std::fma(I.Pu[sk * NU * NM + i + NU * l], I.K[sk * NM * NX + l + NM * j], s);
          }
          Ko[it] = s;
        }
      }
    }
    for (int i = 0; i < NX; ++i) {
      const double dxe = I.dx[static_cast<size_t>(N) * NX + i];
      arm = std::fma(I.q[static_cast<size_t>(N) * NX + i], dxe, arm);
      dxn = std::fma(dxe, dxe, dxn);
    }
    dxn = std::sqrt(dxn);
    dun = std::sqrt(dun);
    double t2 = now();
    tm.qp += t2 - t1;
    // filter line search (SqpSolver::takeStep, FilterLinesearch::acceptStep; accept_kernel)
    double alpha = 1.0, step = 0.0, newMerit = baseMerit, newCost = baseCost, newDyn = baseDyn, newEq = baseEq;
    int stepType = 4;
    const double g0 = std::sqrt(baseDyn + baseEq);
    for (;;) {
      parallelNodes(N + 1, nodeThreads, ws, [&](int k, NodeWs& w) { rolloutNode(m, I, k, alpha, w); });
      double c = 0, dyn = 0, eq = 0;
      for (int k = 0; k <= N; ++k) {
        c += I.ls[4 * k];
        dyn += I.ls[4 * k + 1];
        eq += I.ls[4 * k + 2];
      }
      for (int i = 0; i < NX; ++i) {
        const double df = (1.0 - alpha) * dx0[i];
        dyn = std::fma(df, df, dyn);
      }
      const double g1 = std::sqrt(dyn + eq), armA = alpha * arm;
      bool acc;
      int type;
      if (g1 > st.g_max) {
        acc = g1 < (1.0 - st.gamma_c) * g0;
        type = 1;
      } else if (g1 < st.g_min && g0 < st.g_min && armA < 0.0) {
        acc = c < baseMerit + st.armijo_factor * armA;
        type = 3;
      } else {
        acc = c < baseMerit - st.gamma_c * g0 || g1 < (1.0 - st.gamma_c) * g0;
        type = 2;
      }
      if (acc) {
        step = alpha;
        stepType = type;
        newMerit = newCost = c;
        newDyn = dyn;
        newEq = eq;
        for (size_t i = 0; i < static_cast<size_t>(N + 1) * NX; ++i) I.x[i] = std::fma(alpha, I.dx[i], I.x[i]);
        for (size_t i = 0; i < static_cast<size_t>(N) * NU; ++i) I.u[i] = std::fma(alpha, I.du[i], I.u[i]);
        break;
      }
      const double next = alpha * st.alpha_decay;
      if ((next * dxn < st.delta_tol && next * dun < st.delta_tol) || !(next >= st.alpha_min)) break;  // zero step
      alpha = next;
    }
    tm.ls += now() - t2;
    // checkConvergence (SqpSolver.cpp:583-602)
    int conv = 0;
    if (iter + 1 >= st.sqp_iteration) conv = 1;
    else if (step < st.alpha_min) conv = 2;
    else if (std::fabs(newMerit - baseMerit) < st.cost_tol && std::sqrt(newDyn + newEq) < st.g_min) conv = 3;
    else if (step * dxn < st.delta_tol && step * dun < st.delta_tol) conv = 4;
    if (log) {
      double* Lg = log + 16 * iter;
      const double vals[16] = {baseMerit, baseCost, baseDyn, baseEq, newMerit, newCost, newDyn, newEq, step, static_cast<double>(stepType), step * dxn,
                               step * dun, arm, static_cast<double>(conv), 0, 0};
      std::memcpy(Lg, vals, sizeof(vals));
    }
    ++iter;
    if (conv) break;
  }
  return iter;
}

}  // namespace

extern "C" {

// Batch of `count` instances; arrays as in b200sqp_upload_instances (x / u hold the initial guess on entry, the iterate on return).
// threads: instances in flight; node_threads: workers over the shooting nodes of one instance (CPU-A: threads = 1, node_threads = 4).
// log [count][sqp_iteration][16] and K [count][n_nodes-1][nu*nx] may be null; stage_s [3] = summed seconds of {LQ, QP, line search}.
int orc_fast_wb_sqp_batch(const b200sqp_model_desc* d, int count, int threads, int node_threads, int n_nodes, const double* t_nodes, const uint8_t* events,
                          const double* x0, double* x, double* u, const uint8_t* contact, const double* swing, const double* impact,
                          const double* arm_phase, const double* xref, const b200sqp_settings* st, double* log, int32_t* n_iter, double* K, double* stage_s) {
  static WbDeviceModel m;  // (identical for every call of one process: the benchmark and the tests use one model)
  if (const char* e = makeDeviceModel(*d, m)) {
    std::fprintf(stderr, "orc_fast: %s\n", e);
    return -1;
  }
  const size_t n = static_cast<size_t>(n_nodes);
  std::vector<int> rc(count, 0);
  std::atomic<int> next{0};
  std::vector<Timers> tms(std::max(1, threads));
  auto worker = [&](int tix) {
    std::vector<NodeWs> ws(std::max(1, node_threads));
    Inst I;
    I.alloc(n_nodes - 1, K != nullptr || st->use_feedback_policy);
    for (int i = next++; i < count; i = next++) {
      I.t = t_nodes + i * n;
      I.event = events + i * n;
      I.x0 = x0 + static_cast<size_t>(i) * NX;
      I.x = x + i * n * NX;
      I.u = u + i * (n - 1) * NU;
      I.contact = contact + i * n * 2;
      I.swing = swing + i * n * 6;
      I.impact = impact + i * n * 2;
      I.arm = arm_phase + i * n;
      I.xref = xref + i * n * NX;
      const int it = solveOne(m, I, *st, node_threads, ws, log ? log + static_cast<size_t>(i) * st->sqp_iteration * 16 : nullptr, tms[tix]);
      rc[i] = it < 0 ? -4 : 0;
      if (n_iter) n_iter[i] = it;
      if (K && it >= 0) std::memcpy(K + i * (n - 1) * NU * NX, I.Kre.data(), sizeof(double) * (n - 1) * NU * NX);
    }
  };
  if (threads <= 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto& t : pool) t.join();
  }
  if (stage_s) {
    stage_s[0] = stage_s[1] = stage_s[2] = 0.0;
    for (const auto& t : tms) {
      stage_s[0] += t.lq;
      stage_s[1] += t.qp;
      stage_s[2] += t.ls;
    }
  }
  for (int r : rc)
    if (r) return r;
  return 0;
}

}  // extern "C"
