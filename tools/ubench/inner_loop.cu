// Microbenchmark of the score+grad inner loop (F = 14, RQ(p=2), C = 1) in isolation: a table of rows in shared memory,
// every warp sweeps all rows; no TMA, no barriers, no FK.  Reports FMA-pipe utilisation for scheduling variants.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o inner_loop inner_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

struct P2 {
  float2 v;
  __device__ __forceinline__ P2() {}
  __device__ __forceinline__ P2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit P2(float2 f) : v(f) {}
};
__device__ __forceinline__ P2 padd(P2 x, P2 y) { return P2(__fadd2_rn(x.v, y.v)); }
__device__ __forceinline__ P2 pmul(P2 x, P2 y) { return P2(__fmul2_rn(x.v, y.v)); }
__device__ __forceinline__ P2 pfma(P2 x, P2 y, P2 z) { return P2(__ffma2_rn(x.v, y.v, z.v)); }
__device__ __forceinline__ P2 padd_b(P2 x, float s) { return P2(__fadd2_rn(x.v, make_float2(s, s))); }
__device__ __forceinline__ P2 pmul_b(P2 x, float s) { return P2(__fmul2_rn(x.v, make_float2(s, s))); }
__device__ __forceinline__ P2 pfma_b(float s, P2 y, P2 z) { return P2(__ffma2_rn(make_float2(s, s), y.v, z.v)); }
__device__ __forceinline__ P2 pfma_bb(P2 x, float s, float t) { return P2(__ffma2_rn(x.v, make_float2(s, s), make_float2(t, t))); }
__device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int F = 14, ROW = 16, NROWS = 1024;

__device__ __forceinline__ void load_row(const float* buf, int r, float* rowv) {
  const float4* rp = reinterpret_cast<const float4*>(buf + r * ROW);
#pragma unroll
  for (int i = 0; i < ROW / 4; ++i) { float4 v = rp[i]; rowv[4*i] = v.x; rowv[4*i+1] = v.y; rowv[4*i+2] = v.z; rowv[4*i+3] = v.w; }
}

// VARIANT 0: one row at a time, d kept (the production loop).  1: two rows explicitly interleaved.
// 2: split-form gradient (G += c * (-s), Csum += c; no d kept).  3: two query pairs per lane, d kept.
template <int VARIANT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(float* out, int iters, float c0) {
  extern __shared__ __align__(16) float tab[];
  for (int i = threadIdx.x; i < NROWS * ROW; i += THREADS) tab[i] = 0.001f * (float)((i * 37) % 101) - 0.05f;
  __syncthreads();
  constexpr int NP = (VARIANT == 3) ? 2 : 1;
  P2 x[NP][F], g[NP][F], sc[NP], cs[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
#pragma unroll
    for (int f = 0; f < F; ++f) { x[p][f] = P2(0.01f * threadIdx.x + f + p, 0.02f * threadIdx.x - f); g[p][f] = P2(0.f, 0.f); }
    sc[p] = P2(0.f, 0.f); cs[p] = P2(0.f, 0.f);
  }
  for (int it = 0; it < iters; ++it) {
    if constexpr (VARIANT == 0 || VARIANT == 3) {
#pragma unroll 2
      for (int r = 0; r < NROWS; ++r) {
        float rowv[ROW];
        load_row(tab, r, rowv);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          P2 d[F];
#pragma unroll
          for (int f = 0; f < F; ++f) d[f] = padd_b(x[p][f], rowv[f]);
          P2 a0 = pmul(d[0], d[0]), a1 = pmul(d[1], d[1]);
#pragma unroll
          for (int f = 2; f < F; ++f) { if (f & 1) a1 = pfma(d[f], d[f], a1); else a0 = pfma(d[f], d[f], a0); }
          a0 = padd(a0, a1);
          P2 t = pfma_bb(a0, c0, 1.0f);
          P2 u(rcp(t.v.x), rcp(t.v.y));
          P2 kk = pmul(u, u), coef = pmul(kk, u);
          sc[p] = pfma_b(rowv[F], kk, sc[p]);
          P2 cc = pmul_b(coef, rowv[F]);
#pragma unroll
          for (int f = 0; f < F; ++f) g[p][f] = pfma(cc, d[f], g[p][f]);
        }
      }
    } else if constexpr (VARIANT == 1) {
      for (int r = 0; r < NROWS; r += 2) {
        float ra[ROW], rb[ROW];
        load_row(tab, r, ra);
        load_row(tab, r + 1, rb);
        P2 da[F], db[F];
#pragma unroll
        for (int f = 0; f < F; ++f) { da[f] = padd_b(x[0][f], ra[f]); db[f] = padd_b(x[0][f], rb[f]); }
        P2 a0 = pmul(da[0], da[0]), a1 = pmul(da[1], da[1]), b0 = pmul(db[0], db[0]), b1 = pmul(db[1], db[1]);
#pragma unroll
        for (int f = 2; f < F; ++f) {
          if (f & 1) { a1 = pfma(da[f], da[f], a1); b1 = pfma(db[f], db[f], b1); }
          else { a0 = pfma(da[f], da[f], a0); b0 = pfma(db[f], db[f], b0); }
        }
        a0 = padd(a0, a1); b0 = padd(b0, b1);
        P2 ta = pfma_bb(a0, c0, 1.0f), tb = pfma_bb(b0, c0, 1.0f);
        P2 ua(rcp(ta.v.x), rcp(ta.v.y)), ub(rcp(tb.v.x), rcp(tb.v.y));
        P2 ka = pmul(ua, ua), kb = pmul(ub, ub), ca = pmul(ka, ua), cb = pmul(kb, ub);
        sc[0] = pfma_b(ra[F], ka, sc[0]); sc[0] = pfma_b(rb[F], kb, sc[0]);
        P2 cca = pmul_b(ca, ra[F]), ccb = pmul_b(cb, rb[F]);
#pragma unroll
        for (int f = 0; f < F; ++f) { g[0][f] = pfma(cca, da[f], g[0][f]); g[0][f] = pfma(ccb, db[f], g[0][f]); }
      }
    } else if constexpr (VARIANT == 2) {
#pragma unroll 4
      for (int r = 0; r < NROWS; ++r) {
        float rowv[ROW];
        load_row(tab, r, rowv);
        P2 a0(0.f, 0.f), a1(0.f, 0.f);
#pragma unroll
        for (int f = 0; f < F; ++f) { P2 d = padd_b(x[0][f], rowv[f]); if (f & 1) a1 = pfma(d, d, a1); else a0 = pfma(d, d, a0); }
        a0 = padd(a0, a1);
        P2 t = pfma_bb(a0, c0, 1.0f);
        P2 u(rcp(t.v.x), rcp(t.v.y));
        P2 kk = pmul(u, u), coef = pmul(kk, u);
        sc[0] = pfma_b(rowv[F], kk, sc[0]);
        P2 cc = pmul_b(coef, rowv[F]);
        cs[0] = padd(cs[0], cc);
#pragma unroll
        for (int f = 0; f < F; ++f) g[0][f] = pfma_b(rowv[f], cc, g[0][f]);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
#pragma unroll
    for (int f = 0; f < F; ++f) acc += g[p][f].v.x + g[p][f].v.y;
    acc += sc[p].v.x + sc[p].v.y + cs[p].v.x + cs[p].v.y;
  }
  out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

template <int VARIANT, int THREADS>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  auto kern = k<VARIANT, THREADS>;
  const int smem = NROWS * ROW * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, kern);
  const int iters = 40;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, THREADS, smem>>>(out, 2, 5.0f);
  cudaEventRecord(e0);
  kern<<<148, THREADS, smem>>>(out, iters, 5.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const int np = (VARIANT == 3) ? 2 : 1;
  const double pair_rows = (double)148 * (THREADS / 32) * iters * NROWS * np;  // warp-level (row, query-pair) steps
  const double packed = VARIANT == 2 ? 49.0 : 48.0;
  const double lane_fma = pair_rows * packed * 64;                             // lane-FMA-equivalents
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-44s thr=%3d regs=%3d  %7.3f ms  %6.1f lane-FMA/clk/SM  (%.1f%% of 128)  %6.2f Gpair-evals/s\n", name, THREADS,
         attr.numRegs, ms, lane_fma / (ms * 1e-3) / 148 / (clk * 1e3), 100.0 * lane_fma / (ms * 1e-3) / 148 / (clk * 1e3) / 128.0,
         pair_rows * 64 / (ms * 1e-3) / 1e9);
  cudaFree(out);
}

int main() {
  run<0, 512>("V0 one row, d kept");
  run<0, 384>("V0 one row, d kept");
  run<0, 256>("V0 one row, d kept");
  run<1, 384>("V1 two rows interleaved");
  run<1, 256>("V1 two rows interleaved");
  run<2, 512>("V2 split-form gradient, no d");
  run<2, 384>("V2 split-form gradient, no d");
  run<3, 256>("V3 two query pairs per lane");
  run<3, 384>("V3 two query pairs per lane");
  return 0;
}
