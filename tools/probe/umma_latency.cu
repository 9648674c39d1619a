// tcgen05.mma issue/latency microbenchmark: cycles per MMA for chains of dependent (same accumulator) and independent
// (rotating accumulators) instructions, kind::tf32 (K=8) and kind::f16 (K=16), M=128, several N, A from smem or TMEM.
// Build: nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I diffco_b200/csrc -o tools/probe/umma_latency tools/probe/umma_latency.cu
#include <cstdio>
#include "dc_score_tc.cuh"
namespace dc { long long g_launch_count = 0; }
using namespace dc;

__host__ __device__ constexpr uint32_t idesc_f16(int n) { return umma_idesc_f16(n); }

// kind: 0 tf32, 1 f16.  ts: A from TMEM.  nacc: accumulators rotated.  reps: MMAs issued.
__global__ void __launch_bounds__(128, 2) bench(int kind, int n, int ts, int nacc, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  float* ops = reinterpret_cast<float*>(smem + 128);
  for (int i = threadIdx.x; i < 16384; i += 128) ops[i] = 0.f;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(slot, 256);
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32) {
    const uint32_t a_s = smem_u32(ops), b_s = a_s + 16384;
    const uint64_t ad = umma_desc(a_s, 2048, 128);
    const uint64_t bd = umma_desc(b_s, (uint32_t)n * 16, 128);
    const uint32_t id = kind == 0 ? umma_idesc_tf32(n) : idesc_f16(n);
    const int astride = n <= 64 ? 64 : (n <= 128 ? 128 : 256);   // accumulator column stride
    uint32_t parity = 0;
    for (int rep = 0; rep < 1; ++rep) {
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < reps; ++i) {
          const uint32_t d = tmem + (uint32_t)((i & (nacc - 1)) * astride);
          if (kind == 0) { if (ts) umma_ts(d, tmem + 224, bd, id, 1u); else umma_ss(d, ad, bd, id, 1u); }
          else           { if (ts) umma_f16_ts(d, tmem + 224, bd, id, 1u); else umma_f16_ss(d, ad, bd, id, 1u); }
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait_wd(bar, parity); parity ^= 1;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { atomicMax((unsigned long long*)&out[0], (unsigned long long)(t1 - t0)); atomicMax((unsigned long long*)&out[1], (unsigned long long)(t2 - t0)); }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 65536 + 32768);
  const int reps = 64;
  printf("kind  N  A    nacc   issue cyc/MMA   total cyc/MMA\n");
  for (int kind = 0; kind < 2; ++kind)
    for (int n : {32, 96})
      for (int ts = 0; ts < 2; ++ts)
        for (int nacc : {1}) {
          if (nacc * (n <= 64 ? 64 : (n <= 128 ? 128 : 256)) > 224) continue;
         for (int grid : {1, 296}) {
          cudaMemset(d_out, 0, 16);
          bench<<<grid, 128, 128 + 65536 + 32768>>>(kind, n, ts, nacc, reps, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          printf("%s %4d  %s  %d  grid %3d  %8.1f   %8.1f\n", kind ? "f16 " : "tf32", n, ts ? "tmem" : "smem", nacc, grid, (double)h[0] / reps, (double)h[1] / reps);
         }
        }
  return 0;
}
