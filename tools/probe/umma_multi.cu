// tcgen05.mma issued by SEVERAL warps of one CTA at once: does the per-instruction issue cost (~86 cycles for kind::f16,
// profiles/r01c_umma_issue_cost.txt) parallelise across issuing warps, and may two warps accumulate into the SAME TMEM
// accumulator?  A = B = all ones (f16), so every MMA adds K = 16 to each accumulator element.
// Build: nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I diffco_b200/csrc \
//        -o tools/probe/umma_multi tools/probe/umma_multi.cu
#include <cstdio>
#include <cuda_fp16.h>
#include "dc_score_tc.cuh"
namespace dc { long long g_launch_count = 0; }
using namespace dc;

// nw issuing warps (warp w < nw); same_acc: all accumulate into columns [0, 32), else warp w uses [32 w, 32 w + 32).
// ts: A operand from TMEM (columns 224..231).  reps MMAs per warp.  out[0] = max issue cycles, out[1] = max total cycles.
__global__ void __launch_bounds__(256, 2) bench(int nw, int same_acc, int ts, int n, int reps, long long* out, float* dval) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);  // [8]
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  __half* ops = reinterpret_cast<__half*>(smem + 128);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) ops[i] = __float2half(1.0f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(slot, 256);
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  // zero the accumulators, write the TMEM A operand (f16x2 ones in 8 columns)
  if (warp < 4) {
    uint32_t z[16], o[8];
    for (int i = 0; i < 16; ++i) z[i] = 0;
    for (int i = 0; i < 8; ++i) o[i] = 0x3c003c00u;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 224; c += 16) tmem_st16(tl + c, z);
    tmem_st8(tl + 224, o);
    tmem_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const long long t0 = clock64();
  long long t1 = t0, t2 = t0;
  if (warp < nw) {
    const uint32_t a_s = smem_u32(ops), b_s = a_s + 16384;
    const uint64_t ad = umma_desc(a_s, 2048, 128);
    const uint64_t bd = umma_desc(b_s, (uint32_t)n * 16, 128);
    const uint32_t id = umma_idesc_f16(n);
    const uint32_t d = tmem + (same_acc ? 0u : (uint32_t)(warp * 32));
    if (elect_one()) {
#pragma unroll 4
      for (int i = 0; i < reps; ++i) {
        if (ts) umma_f16_ts(d, tmem + 224, bd, id, 1u); else umma_f16_ss(d, ad, bd, id, 1u);
      }
    }
    __syncwarp();
    t1 = clock64();
    if (elect_one()) umma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait_wd(&bar[warp], 0);
    t2 = clock64();
    if (lane == 0) {
      atomicMax((unsigned long long*)&out[0], (unsigned long long)(t1 - t0));
      atomicMax((unsigned long long*)&out[1], (unsigned long long)(t2 - t0));
    }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp < 4 && blockIdx.x == 0) {
    uint32_t r[16];
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    for (int w = 0; w < 4; ++w) {
      tmem_ld16(tl + w * 32, r);
      tmem_wait_ld();
      dval[(warp * 32 + lane) * 4 + w] = __uint_as_float(r[3]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d_out; float* d_val;
  cudaMalloc(&d_out, 16); cudaMalloc(&d_val, 128 * 4 * 4);
  const int smem = 128 + 32768 + 16384;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 64, n = 32;
  printf("A     warps same_acc grid  issue cyc/MMA(per warp)  total cyc/MMA(all warps)   D check (expect %d per issuing warp)\n", 16 * reps);
  for (int ts = 0; ts < 2; ++ts)
    for (int nw : {1, 2, 3, 4})
      for (int same : {0, 1}) {
        if (nw == 1 && same) continue;
        for (int grid : {1, 296}) {
          cudaMemset(d_out, 0, 16); cudaMemset(d_val, 0, 128 * 4 * 4);
          bench<<<grid, 256, smem>>>(nw, same, ts, n, reps, d_out, d_val);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; float v[512];
          cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(v, d_val, sizeof(v), cudaMemcpyDeviceToHost);
          float mn[4] = {1e30f, 1e30f, 1e30f, 1e30f}, mx[4] = {0, 0, 0, 0};
          for (int r = 0; r < 128; ++r) for (int w = 0; w < 4; ++w) { mn[w] = fminf(mn[w], v[r * 4 + w]); mx[w] = fmaxf(mx[w], v[r * 4 + w]); }
          printf("%s  %d     %d        %3d   %8.1f                 %8.1f                    acc0 [%g,%g] acc1 [%g,%g] acc2 [%g,%g] acc3 [%g,%g]\n",
                 ts ? "tmem" : "smem", nw, same, grid, (double)h[0] / reps, (double)h[1] / (reps * nw), mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mn[3], mx[3]);
        }
      }
  return 0;
}
