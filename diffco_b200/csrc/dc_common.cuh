// Shared device/host helpers for the diffco_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/diffco_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "diffco_b200 kernels are written for sm_100a (B200) only"
#endif

namespace dc {

extern long long g_launch_count;  // host-side counter, bumped by every launch wrapper (dc_launch_count()).

#define DC_CUDA_OK(expr)                         \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) return DC_ERR_CUDA;   \
  } while (0)

#define DC_LAUNCH_CHECK()                        \
  do {                                           \
    ++dc::g_launch_count;                        \
    if (cudaPeekAtLastError() != cudaSuccess) {  \
      (void)cudaGetLastError();                  \
      return DC_ERR_CUDA;                        \
    }                                            \
  } while (0)

constexpr int kWarp = 32;

// Per-device one-time initialisation (cudaFuncSetAttribute is a per-device attribute; one process may score on several
// GPUs).  One instance per kernel instantiation; doing the initialisation twice from racing threads is harmless.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  // returns the current device ordinal (< 64) in *dev and whether its bit is still clear; -1 on error
  int pending(int* dev) {
    if (cudaGetDevice(dev) != cudaSuccess) {
      (void)cudaGetLastError();
      return -1;
    }
    if (*dev < 0 || *dev >= 64) return 1;  // never cached: always initialise
    return (done.load(std::memory_order_acquire) >> *dev) & 1ull ? 0 : 1;
  }
  void mark(int dev) {
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  }
};
#define DC_SET_FUNC_ATTR_ONCE(kern, attr, value)                         \
  do {                                                                   \
    static dc::PerDeviceOnce _once;                                      \
    int _dev = 0;                                                        \
    const int _p = _once.pending(&_dev);                                 \
    if (_p < 0) return DC_ERR_NO_DEVICE;                                 \
    if (_p > 0) {                                                        \
      DC_CUDA_OK(cudaFuncSetAttribute(kern, attr, value));               \
      _once.mark(_dev);                                                  \
    }                                                                    \
  } while (0)

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ----------------------------------------------------------------------------------------------
// Packed pairs.  fp32 pairs map onto Blackwell's 2-wide FP32 instructions (FADD2/FMUL2/FFMA2 in SASS,
// add/mul/fma.f32x2 in PTX): one issue slot, two lanes of work.  The *_b forms take a scalar that the
// instruction broadcasts to both halves (SASS operand Rx.F32).  Measured on B200 (tools/ubench/fma_pipes.cu):
// a loop of packed ops alone sustains ~124 of the 128 lane-FMAs/clk/SM, scalar FFMA alone ~119, but a MIX of
// packed and scalar FMA-pipe instructions drops to ~103 — so the inner loops keep everything packed.
// ----------------------------------------------------------------------------------------------
struct P2 {
  float2 v;
  __device__ __forceinline__ P2() {}
  __device__ __forceinline__ P2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit P2(float2 f) : v(f) {}
  __device__ __forceinline__ float lo() const { return v.x; }
  __device__ __forceinline__ float hi() const { return v.y; }
};
__device__ __forceinline__ P2 padd(P2 x, P2 y) { return P2(__fadd2_rn(x.v, y.v)); }
__device__ __forceinline__ P2 pmul(P2 x, P2 y) { return P2(__fmul2_rn(x.v, y.v)); }
__device__ __forceinline__ P2 pfma(P2 x, P2 y, P2 z) { return P2(__ffma2_rn(x.v, y.v, z.v)); }
__device__ __forceinline__ P2 padd_b(P2 x, float s) { return P2(__fadd2_rn(x.v, make_float2(s, s))); }
__device__ __forceinline__ P2 pmul_b(P2 x, float s) { return P2(__fmul2_rn(x.v, make_float2(s, s))); }
__device__ __forceinline__ P2 pfma_b(float s, P2 y, P2 z) { return P2(__ffma2_rn(make_float2(s, s), y.v, z.v)); }
__device__ __forceinline__ P2 pfma_bb(P2 x, float s, float t) {  // x * s + t
  return P2(__ffma2_rn(x.v, make_float2(s, s), make_float2(t, t)));
}

// ----------------------------------------------------------------------------------------------
// Scalar math with the precision each dtype needs.  fp32: MUFU.RCP / MUFU.RSQ approximations
// (<= 2^-22 relative error, far inside the 1e-5 parity gate); fp64: IEEE division / sqrt.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }

template <typename T>
struct Tiny;
template <>
struct Tiny<float> {
  static constexpr float v = 1e-30f;
};
template <>
struct Tiny<double> {
  static constexpr double v = 1e-300;
};

__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }

// ----------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) — the support table is a contiguous array of
// 16-byte-aligned rows, so a chunk of rows is one linear bulk copy; no tensor map is needed.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared::cta bulk copy, completion signalled on `bar` (complete_tx).  bytes % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace dc
