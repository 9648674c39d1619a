// Microbenchmark: FP32 pipe throughput on sm_100a for scalar FFMA, packed FFMA2 and mixes (lane-FMAs per clock per SM).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_pipes fma_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float seed) {
  // 8 independent chains per thread
  float2 a[8];
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); s[i] = seed * i; }
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (MODE == 0) {  // scalar FFMA only: 16 per r
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      } else if (MODE == 1) {  // packed FFMA2 only: 8 per r (16 lane-FMAs)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
      } else if (MODE == 2) {  // 8 FFMA2 + 8 scalar FFMA (32 lane-FMA-equivalents... 16 + 8 = 24 lane FMAs)
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = __ffma2_rn(a[i], m, c); s[i] = fmaf(s[i], m.x, c.y); }
      } else if (MODE == 3) {  // 8 FFMA2 + 4 scalar
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = __ffma2_rn(a[i], m, c); if (i < 4) s[i] = fmaf(s[i], m.x, c.y); }
      } else if (MODE == 4) {  // FADD2 only
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fadd2_rn(a[i], c);
      } else if (MODE == 5) {  // 8 FFMA2 + 2 MUFU.RCP
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[0]));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[1]));
      } else if (MODE == 6) {  // FFMA2 with scalar-broadcast operand (like the gradient update)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(make_float2(s[i & 1], s[i & 1]), m, a[i]);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double lane_fma_per_iter, int threads) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 1000, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double total = (double)148 * threads * iters * 4 * lane_fma_per_iter;
  printf("%-34s threads=%4d  %8.3f ms  %7.2f lane-FMA/clk/SM (at %d MHz nominal)  %6.2f TFLOP/s\n", name, threads, ms,
         total / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, 2 * total / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("scalar FFMA x16", 16, threads);
    run<1>("FFMA2 x8", 16, threads);
    run<2>("FFMA2 x8 + FFMA x8", 24, threads);
    run<3>("FFMA2 x8 + FFMA x4", 20, threads);
    run<4>("FADD2 x8", 16, threads);
    run<5>("FFMA2 x8 + MUFU x2", 16, threads);
    run<6>("FFMA2 x8 bcast-scalar operand", 16, threads);
  }
  return 0;
}
