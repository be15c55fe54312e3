// Micro-probe (run under gpurun): fp64 FMA latency / throughput per SM, barrier and shared-memory round trips on
// sm_100a.  Explains the cost structure of the single-CTA fp64 kernels (k_lm_step, k_assemble, k_pair_setup).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double* out, int n, long long* cyc) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  out[3] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dfma_tput(double* out, int n, long long* cyc) {
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = out[k & 3] + k + threadIdx.x;
  const double b = out[1], c = out[2];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[4 + threadIdx.x % 4] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ffma_tput(float* out, int n, long long* cyc) {
  float a[8];
  for (int k = 0; k < 8; ++k) a[k] = out[k & 3] + k + threadIdx.x;
  const float b = out[1], c = out[2];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], b, c);
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[4 + threadIdx.x % 4] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void barrier_cost(int n, long long* cyc) {
  __shared__ double sh[256];
  long long t0 = clock64();
  double v = threadIdx.x;
  for (int i = 0; i < n; ++i) {
    sh[threadIdx.x] = v;
    __syncthreads();
    v += sh[(threadIdx.x + 1) & 255];
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0 + (v == 1e300);
}
__global__ void rcp_cost(double* out, int n, long long* cyc) {
  double d = out[0] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double inv = (double)__frcp_rn((float)d);
    inv = inv * (2.0 - d * inv);
    inv = inv * (2.0 - d * inv);
    d = inv + 1.0;
  }
  long long t1 = clock64();
  out[3] = d;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void div_cost(double* out, int n, long long* cyc) {
  double d = out[0] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) d = 1.0 / d + 1.0;
  long long t1 = clock64();
  out[3] = d;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d;
  float* f;
  long long* c;
  cudaMalloc(&d, 64 * 8);
  cudaMalloc(&f, 64 * 4);
  cudaMallocManaged(&c, 64);
  double h[8] = {1.0, 0.999999, 1e-9, 0, 0, 0, 0, 0};
  float hf[8] = {1.0f, 0.999999f, 1e-9f, 0, 0, 0, 0, 0};
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemcpy(f, hf, sizeof(hf), cudaMemcpyHostToDevice);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    dfma_chain<<<1, 32>>>(d, n, c); cudaDeviceSynchronize();
    if (rep) printf("dfma dependent chain: %.1f cycles per fma (1 warp)\n", (double)c[0] / n);
    for (int threads : {32, 128, 256, 512, 1024}) {
      dfma_tput<<<1, threads>>>(d, n, c); cudaDeviceSynchronize();
      if (rep) printf("dfma throughput %4d threads: %.2f fma/clk/SM\n", threads, (double)n * 8 * threads / c[0]);
    }
    for (int threads : {256, 1024}) {
      ffma_tput<<<1, threads>>>(f, n, c); cudaDeviceSynchronize();
      if (rep) printf("ffma throughput %4d threads: %.2f fma/clk/SM\n", threads, (double)n * 8 * threads / c[0]);
    }
    barrier_cost<<<1, 256>>>(n, c); cudaDeviceSynchronize();
    if (rep) printf("sts + barrier + lds + barrier (256 thr): %.1f cycles\n", (double)c[0] / n);
    rcp_cost<<<1, 32>>>(d, n, c); cudaDeviceSynchronize();
    if (rep) printf("rcp64 (frcp + 2 newton) chain: %.1f cycles\n", (double)c[0] / n);
    div_cost<<<1, 32>>>(d, n, c); cudaDeviceSynchronize();
    if (rep) printf("ieee 1/d chain: %.1f cycles\n", (double)c[0] / n);
  }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock rate %d kHz, err %s\n", clk, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
