// C ABI implementation (include/b200sqp.h).  Host C++ orchestration + kernel launches; no CPU fallback.
#include "../../include/b200sqp.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "riccati.cuh"
#include "riccati_cluster.cuh"
#include "riccati_wb.cuh"
#ifdef B200SQP_WITH_WB
#include "wb_solver.cuh"
#include "wb_builder.cuh"
#include "cen_dynamics.cuh"
#include "wb_torque.cuh"
#include "cen_kernels.cuh"
#include "cen_host.cuh"
#endif

namespace {
thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? B200SQP_ENOMEM : B200SQP_ENODEV, "%s: %s", #expr, \
                                       cudaGetErrorString(e_));                                    \
  } while (0)

int select_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(B200SQP_ENODEV, "no CUDA device available (%s); b200sqp has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(B200SQP_EINVAL, "device %d out of range (%d devices)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  return 0;
}

cudaError_t set_riccati_smem_attributes(int nx, int numax) {
  const int one = static_cast<int>(b200sqp::riccati_smem_doubles(nx, numax) * sizeof(double));
  const int clu = static_cast<int>(b200sqp::riccati_cluster_smem_doubles(nx, numax) * sizeof(double));
  cudaError_t e = cudaFuncSetAttribute(b200sqp::riccati_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, one);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(b200sqp::riccati_cluster_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, clu);
  if (nx == 58 && numax == 23) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(b200sqp::ricwb::riccati_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(b200sqp::ricwb::BWD_SMEM_BYTES));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(b200sqp::ricwb::riccati_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(b200sqp::ricwb::FWD_SMEM_BYTES));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(b200sqp::riccati_kernel<58, 23>, cudaFuncAttributeMaxDynamicSharedMemorySize, one);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(b200sqp::riccati_cluster_kernel<58, 23>, cudaFuncAttributeMaxDynamicSharedMemorySize, clu);
  }
  return e;
}

// K2 dispatch: the cluster variant (RC SMs per instance) for small batches; for the whole-body sizes the two-kernel form of riccati_wb.cuh
// (backward factorisation with a helper warp + separate forward substitution); the generic one-CTA kernel for every other size.
// B200SQP_NO_CLUSTER=1 skips the cluster variant, B200SQP_K2_LEGACY=1 forces the generic kernel (tests exercise all three).
// Returns the number of kernels launched.
int launch_riccati(const b200sqp::QpDeviceView& v, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const char* no = std::getenv("B200SQP_NO_CLUSTER");
  // measured on B200: the cluster variant wins while at most ~half of the SMs are taken by clusters
  const bool cluster = !(no && no[0] == '1') && v.B * b200sqp::RC * 2 <= sms;
  const bool wb = (v.nx == 58 && v.numax == 23);   // the whole-body sizes have a compile-time instantiation
  if (cluster) {
    const size_t smem = b200sqp::riccati_cluster_smem_doubles(v.nx, v.numax) * sizeof(double);
    if (wb) b200sqp::riccati_cluster_kernel<58, 23><<<v.B * b200sqp::RC, 256, smem, st>>>(v);
    else b200sqp::riccati_cluster_kernel<0, 0><<<v.B * b200sqp::RC, 256, smem, st>>>(v);
  } else {
    const char* legacy = std::getenv("B200SQP_K2_LEGACY");
    if (wb && !(legacy && legacy[0] == '1')) {
      b200sqp::ricwb::riccati_bwd_kernel<<<v.B, b200sqp::ricwb::BWD_THREADS, b200sqp::ricwb::BWD_SMEM_BYTES, st>>>(v);
      b200sqp::ricwb::riccati_fwd_kernel<<<v.B, b200sqp::ricwb::FWD_THREADS, b200sqp::ricwb::FWD_SMEM_BYTES, st>>>(v);
      return 2;
    }
    const size_t smem = b200sqp::riccati_smem_doubles(v.nx, v.numax) * sizeof(double);
    if (wb) b200sqp::riccati_kernel<58, 23><<<v.B, 256, smem, st>>>(v);
    else b200sqp::riccati_kernel<0, 0><<<v.B, 256, smem, st>>>(v);
  }
  return 1;
}
}  // namespace

struct b200sqp_qp_t {
  int device = 0, B = 0, N = 0, nx = 0, numax = 0;
  double *A = nullptr, *Bm = nullptr, *b = nullptr, *Q = nullptr, *S = nullptr, *R = nullptr, *q = nullptr, *r = nullptr, *dx0 = nullptr;
  int* nu = nullptr;
  bool hasNu = false;
  double *K = nullptr, *kff = nullptr, *P = nullptr, *p = nullptr, *dx = nullptr, *du = nullptr;
  int* status = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool uploaded = false, solved = false, keptP = false;
};

extern "C" {

const char* b200sqp_last_error(void) { return g_err.c_str(); }
const char* b200sqp_version(void) { return "b200sqp 0.1 (sm_100a, fp64)"; }

int b200sqp_qp_create(int device, int batch, int N, int nx, int nu_max, b200sqp_qp* out) {
  if (!out || batch <= 0 || N <= 0 || nx <= 0 || nx > 64 || nu_max <= 0 || nu_max > 32)
    return fail(B200SQP_EINVAL, "qp_create: need batch,N > 0, 0 < nx <= 64 and 0 < nu_max <= 32 (got %d,%d,%d,%d)", batch, N, nx, nu_max);
  if (int rc = select_device(device)) return rc;
  const size_t smem = b200sqp::riccati_smem_doubles(nx, nu_max) * sizeof(double);
  if (smem > 227 * 1024) return fail(B200SQP_EINVAL, "qp_create: nx=%d nu_max=%d needs %zu B shared memory (> 227 KB)", nx, nu_max, smem);
  b200sqp_qp_t* qp = new (std::nothrow) b200sqp_qp_t;
  if (!qp) return fail(B200SQP_ENOMEM, "host allocation failed");
  qp->device = device;
  qp->B = batch;
  qp->N = N;
  qp->nx = nx;
  qp->numax = nu_max;
  const size_t bn = static_cast<size_t>(batch) * N, bn1 = static_cast<size_t>(batch) * (N + 1);
  auto alloc = [&](double** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(double)); };
  cudaError_t e = cudaSuccess;
  auto A_ = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A_(alloc(&qp->A, bn * nx * nx));
  A_(alloc(&qp->Bm, bn * nx * nu_max));
  A_(alloc(&qp->b, bn * nx));
  A_(alloc(&qp->Q, bn1 * nx * nx));
  A_(alloc(&qp->S, bn * nu_max * nx));
  A_(alloc(&qp->R, bn * nu_max * nu_max));
  A_(alloc(&qp->q, bn1 * nx));
  A_(alloc(&qp->r, bn * nu_max));
  A_(alloc(&qp->dx0, static_cast<size_t>(batch) * nx));
  A_(cudaMalloc(reinterpret_cast<void**>(&qp->nu), bn * sizeof(int)));
  A_(alloc(&qp->K, bn * nu_max * nx));
  A_(alloc(&qp->kff, bn * nu_max));
  A_(alloc(&qp->P, bn1 * nx * nx));
  A_(alloc(&qp->p, bn1 * nx));
  A_(alloc(&qp->dx, bn1 * nx));
  A_(alloc(&qp->du, bn * nu_max));
  A_(cudaMalloc(reinterpret_cast<void**>(&qp->status), batch * sizeof(int)));
  A_(cudaEventCreate(&qp->ev0));
  A_(cudaEventCreate(&qp->ev1));
  if (e == cudaSuccess)
    e = set_riccati_smem_attributes(nx, nu_max);
  if (e != cudaSuccess) {
    b200sqp_qp_destroy(qp);
    return fail(e == cudaErrorMemoryAllocation ? B200SQP_ENOMEM : B200SQP_ENODEV, "qp_create: %s", cudaGetErrorString(e));
  }
  *out = qp;
  return 0;
}

void b200sqp_qp_destroy(b200sqp_qp qp) {
  if (!qp) return;
  cudaSetDevice(qp->device);
  for (double* p : {qp->A, qp->Bm, qp->b, qp->Q, qp->S, qp->R, qp->q, qp->r, qp->dx0, qp->K, qp->kff, qp->P, qp->p, qp->dx, qp->du})
    if (p) cudaFree(p);
  if (qp->nu) cudaFree(qp->nu);
  if (qp->status) cudaFree(qp->status);
  if (qp->ev0) cudaEventDestroy(qp->ev0);
  if (qp->ev1) cudaEventDestroy(qp->ev1);
  delete qp;
}

int b200sqp_qp_upload(b200sqp_qp qp, const double* A, const double* B, const double* b, const double* Q, const double* S,
                      const double* R, const double* q, const double* r, const int32_t* nu, const double* dx0) {
  if (!qp || !A || !B || !b || !Q || !S || !R || !q || !r || !dx0) return fail(B200SQP_EINVAL, "qp_upload: null argument");
  CUDA_TRY(cudaSetDevice(qp->device));
  const size_t bn = static_cast<size_t>(qp->B) * qp->N, bn1 = static_cast<size_t>(qp->B) * (qp->N + 1);
  const int nx = qp->nx, nm = qp->numax;
  auto up = [&](double* d, const double* h, size_t n) { return cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice); };
  CUDA_TRY(up(qp->A, A, bn * nx * nx));
  CUDA_TRY(up(qp->Bm, B, bn * nx * nm));
  CUDA_TRY(up(qp->b, b, bn * nx));
  CUDA_TRY(up(qp->Q, Q, bn1 * nx * nx));
  CUDA_TRY(up(qp->S, S, bn * nm * nx));
  CUDA_TRY(up(qp->R, R, bn * nm * nm));
  CUDA_TRY(up(qp->q, q, bn1 * nx));
  CUDA_TRY(up(qp->r, r, bn * nm));
  CUDA_TRY(up(qp->dx0, dx0, static_cast<size_t>(qp->B) * nx));
  qp->hasNu = nu != nullptr;
  if (nu) {
    for (size_t i = 0; i < bn; ++i)
      if (nu[i] < 0 || nu[i] > nm) return fail(B200SQP_EINVAL, "qp_upload: nu[%zu]=%d outside [0,%d]", i, nu[i], nm);
    CUDA_TRY(cudaMemcpy(qp->nu, nu, bn * sizeof(int), cudaMemcpyHostToDevice));
  }
  qp->uploaded = true;
  qp->solved = false;
  return 0;
}

int b200sqp_qp_solve(b200sqp_qp qp, double reg_prim, int keep_P, void* stream) {
  if (!qp) return fail(B200SQP_EINVAL, "qp_solve: null handle");
  if (!qp->uploaded) return fail(B200SQP_ESTATE, "qp_solve: no problem uploaded");
  CUDA_TRY(cudaSetDevice(qp->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bn = static_cast<size_t>(qp->B) * qp->N;
  b200sqp::QpDeviceView v{};
  v.B = qp->B;
  v.N = qp->N;
  v.nx = qp->nx;
  v.numax = qp->numax;
  v.A = qp->A;
  v.Bm = qp->Bm;
  v.b = qp->b;
  v.Q = qp->Q;
  v.S = qp->S;
  v.R = qp->R;
  v.q = qp->q;
  v.r = qp->r;
  v.dx0 = qp->dx0;
  v.nu = qp->hasNu ? qp->nu : nullptr;
  v.K = qp->K;
  v.kff = qp->kff;
  v.P = qp->P;
  v.p = qp->p;
  v.dx = qp->dx;
  v.du = qp->du;
  v.status = qp->status;
  v.keepP = keep_P;
  v.reg = reg_prim;
  CUDA_TRY(cudaEventRecord(qp->ev0, st));
  CUDA_TRY(cudaMemsetAsync(qp->K, 0, bn * qp->numax * qp->nx * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(qp->kff, 0, bn * qp->numax * sizeof(double), st));
  launch_riccati(v, st);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(qp->ev1, st));
  qp->solved = true;
  qp->keptP = keep_P != 0;
  return 0;
}

int b200sqp_qp_download(b200sqp_qp qp, double* dx, double* du, double* K, double* k, double* P, double* p, int32_t* status) {
  if (!qp) return fail(B200SQP_EINVAL, "qp_download: null handle");
  if (!qp->solved) return fail(B200SQP_ESTATE, "qp_download: nothing solved yet");
  if ((P || p) && !qp->keptP) return fail(B200SQP_ESTATE, "qp_download: cost-to-go requested but solve ran with keep_P = 0");
  CUDA_TRY(cudaSetDevice(qp->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t bn = static_cast<size_t>(qp->B) * qp->N, bn1 = static_cast<size_t>(qp->B) * (qp->N + 1);
  const int nx = qp->nx, nm = qp->numax;
  auto dn = [&](double* h, const double* d, size_t n) { return h ? cudaMemcpy(h, d, n * sizeof(double), cudaMemcpyDeviceToHost) : cudaSuccess; };
  CUDA_TRY(dn(dx, qp->dx, bn1 * nx));
  CUDA_TRY(dn(du, qp->du, bn * nm));
  CUDA_TRY(dn(K, qp->K, bn * nm * nx));
  CUDA_TRY(dn(k, qp->kff, bn * nm));
  CUDA_TRY(dn(P, qp->P, bn1 * nx * nx));
  CUDA_TRY(dn(p, qp->p, bn1 * nx));
  std::vector<int> st(qp->B);
  CUDA_TRY(cudaMemcpy(st.data(), qp->status, qp->B * sizeof(int), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < qp->B; ++i) {
    if (status) status[i] = st[i];
    bad += st[i] != 0;
  }
  if (bad) return fail(B200SQP_EQP, "[b200sqp] Failed to solve QP for %d of %d instances", bad, qp->B);
  return 0;
}

int b200sqp_qp_last_ms(b200sqp_qp qp, float* ms) {
  if (!qp || !ms) return fail(B200SQP_EINVAL, "qp_last_ms: null argument");
  if (!qp->solved) return fail(B200SQP_ESTATE, "qp_last_ms: nothing solved yet");
  CUDA_TRY(cudaSetDevice(qp->device));
  CUDA_TRY(cudaEventSynchronize(qp->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms, qp->ev0, qp->ev1));
  return 0;
}

void b200sqp_default_settings(b200sqp_settings* s) {
  if (!s) return;
  // sqp::Settings defaults (SqpSettings.h:42-86) overridden by G1 task.info:77-94
  s->sqp_iteration = 1;
  s->delta_tol = 1e-4;
  s->cost_tol = 1e-4;
  s->alpha_decay = 0.5;
  s->alpha_min = 1e-4;
  s->gamma_c = 1e-6;
  s->g_max = 1e-2;
  s->g_min = 1e-6;
  s->armijo_factor = 1e-4;
  s->reg_prim = 1e-12;
  s->use_feedback_policy = 0;
  s->global_step = 0;
  s->create_value_function = 0;
}

}  // extern "C"

#ifdef B200SQP_WITH_WB
#include "wb_capi.inc"
#endif
