// Micro-probe (run under gpurun): the fp32 CUDA-core ceiling of this B200, measured -- the roof the issue-bound fused
// linearise is reported against (SURVEY section 0 / VERDICT r01 item 8: "the builder must measure it").
//   * FFMA  : one warp-wide fused multiply-add per issue slot
//   * FFMA2 : sm_100's packed pair (PTX fma.rn.f32x2): two fused multiply-adds per thread per issue slot
//   * mixed : FFMA2 interleaved with integer ALU work, to see whether the issue slots FFMA2 frees are usable
// Per-SM rates from one CTA with clock64(); the whole-chip rate from a grid of 148 x 4 CTAs with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_probe tools/fp32_probe.cu && tools/fp32_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>  // 0 FFMA, 1 FFMA2, 2 FFMA2 + IMAD interleaved, 3 FFMA + IMAD interleaved
__global__ void tput(float* out, int n, long long* cyc) {
  float a[16];
  unsigned long long p[8];
  int z[8];
  for (int k = 0; k < 16; ++k) a[k] = out[k & 3] + k + threadIdx.x;
  for (int k = 0; k < 8; ++k) p[k] = pack(a[2 * k], a[2 * k + 1]), z[k] = threadIdx.x + k;
  const float b = out[1], c = out[2];
  const unsigned long long pb = pack(b, b), pcc = pack(c, c);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0 || MODE == 3) {
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], b, c);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) p[k] = ffma2(p[k], pb, pcc);
    }
    if (MODE == 2 || MODE == 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k) z[k] = z[k] * 3 + i;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
  for (int k = 0; k < 16; ++k) s += a[k];
  for (int k = 0; k < 8; ++k) s += __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32)) + (float)z[k];
  out[4 + (blockIdx.x * blockDim.x + threadIdx.x) % 4] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* f;
  long long* c;
  cudaMalloc(&f, 64 * 4);
  cudaMallocManaged(&c, 64);
  float hf[8] = {1.0f, 0.999999f, 1e-9f, 0, 0, 0, 0, 0};
  cudaMemcpy(f, hf, sizeof(hf), cudaMemcpyHostToDevice);
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int n = 4096;
  const char* names[4] = {"FFMA", "FFMA2 (fma.rn.f32x2)", "FFMA2 + 8 IMAD per 8 FFMA2", "FFMA + 8 IMAD per 16 FFMA"};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 4; ++mode)
      for (int threads : {128, 256, 512, 1024}) {
        if (mode == 0) tput<0><<<1, threads>>>(f, n, c);
        if (mode == 1) tput<1><<<1, threads>>>(f, n, c);
        if (mode == 2) tput<2><<<1, threads>>>(f, n, c);
        if (mode == 3) tput<3><<<1, threads>>>(f, n, c);
        cudaDeviceSynchronize();
        if (rep) printf("%-28s %4d threads: %6.1f fma/clk/SM (%5.2f warp-instr issue slots/clk/SM incl. the integer ones)\n", names[mode], threads,
                        (double)n * 16 * threads / c[0],
                        (double)n * ((mode == 0 || mode == 3 ? 16 : 8) + (mode >= 2 ? 8 : 0)) * threads / 32.0 / c[0]);
      }
  // whole chip, CUDA events
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    const int reps = 20, nn = 1 << 15;
    for (int w = 0; w < 2; ++w) {
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) tput<0><<<sms * 4, 512>>>(f, nn, c);
        else tput<1><<<sms * 4, 512>>>(f, nn, c);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * nn * 512.0 * sms * 4 * reps;
    printf("whole chip %-22s: %.1f TFLOP/s fp32 (%d SMs, grid %d x 512, %d launches in %.3f ms)\n", names[mode], flops / (ms * 1e-3) / 1e12, sms,
           sms * 4, reps, ms);
  }
  printf("clock rate %d kHz, err %s\n", clk, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
