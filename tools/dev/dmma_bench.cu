// Development aid: fp64 tensor-path (DMMA m8n8k4) issue rate and dependent latency on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dev/dmma_bench tools/dev/dmma_bench.cu ; gpurun -- tools/dev/dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void dmma_chain(int iters, double* out, long long* cyc) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c0[i] = c1[i] = threadIdx.x * 1e-9 + i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// DFMA reference: per-thread dependent chains
template <int CH>
__global__ void dfma_chain(int iters, double* out, long long* cyc) {
  double c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-9 + i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CH>
void run(int warps, bool mma) {
  double* out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) {
    if (mma) dmma_chain<CH><<<148, warps * 32>>>(iters, out, cyc);
    else dfma_chain<CH><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (double(iters) * CH);
  if (mma)
    printf("DMMA chains/warp %d warps/SM %2d: %.2f cycles per DMMA per warp -> %.1f cycles per DMMA per SM, %.1f FMA/clk/SM\n", CH, warps, per, per / warps,
           256.0 * warps / per);
  else
    printf("DFMA chains/thread %d warps/SM %2d: %.2f cycles per DFMA per warp -> %.1f FMA/clk/SM\n", CH, warps, per, 32.0 * warps / per);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4, 8, 9, 16}) {
    run<1>(w, true);
    run<2>(w, true);
    run<4>(w, true);
    run<8>(w, true);
  }
  for (int w : {1, 4, 8, 16}) {
    run<1>(w, false);
    run<4>(w, false);
  }
  return 0;
}
