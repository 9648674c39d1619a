// Fused collision-score kernel, tensor-core form (tcgen05 + TMEM + bulk TMA; RQKernel(p = 2), one class, fp32 I/O, F <= 14).
//
//   score[b] = sum_n w_n k(rho_bn)        g_x[b] = -2 gamma sum_n w_n u^3 (x_b - s_n)        rho_bn = |x_b - s_n|^2
//
// Same contract as score_tq_kernel (dc_score_tq.cuh) for diffco/kernel_perceptrons.py:362-370 (DiffCo.score with
// diffco/kernel.py:17-29 and a diffco/model.py feature map) and its autograd backward, but the two contractions run on
// the 5th-generation tensor cores instead of the FP32 pipe:
//
//   GEMM1  rho[128 x 96]  = A[128 x 48] . B1[96 x 48]^T     (kind::f16, K = 48 = 3 x 16 slots, fp32 accumulate in TMEM)
//   GEMM2  G  [128 x 32] += CC[128 x 16] . B2[32 x 16]^T    per 8 support vectors, CC read from TMEM
//
// Every fp32 quantity enters as a sum of 11-bit terms (x = xh + xl, s = sh + sl, cc = ch + cl) placed in separate K slots:
//   rho = sum_f xh(-2sh) + xl(-2sh) + xh(-2s)_l  +  |s|^2 (3 terms x 1) + |x|^2 (3 terms x 1)          48 slots
//   G[:, 0:16]  = sum_n (ch + cl) [sh | 1]          G[:, 16:32] = sum_n ch sl        g_x = -2 gamma (x G[14] - G[f] - G[16+f])
// i.e. products carry ~22 bits and are accumulated in fp32.  Power-of-two scales (features by Sx, weights by Sw, chosen
// at pack time) keep every term inside fp16's exponent range and are undone exactly in the epilogue.
// The expansion |x|^2 + |s|^2 - 2 x.s leaves an ABSOLUTE error of ~1e-5 (scaled by the feature magnitudes) in rho, which
// matters only where k'(rho) is large, i.e. for the few pairs with small rho: those (rho below a per-query threshold
// derived from the error bound) are re-evaluated exactly with direct differences on the FP32 pipe and removed from the
// tensor-core sums.  Every other pair's error is below tol_pair (2e-7) of its weight.
//
// One tcgen05.mma occupies the tensor pipe for >= ~85 cycles whatever its N (profiles/r01c_umma_issue_cost.txt), so the
// instruction count per support vector is what bounds this kernel: 3 instructions per 96 supports for GEMM1 and one
// per 8 supports for GEMM2 (the K-slot packing above exists to get there).
//
// CTA = 256 query threads (8 warps: TMEM lane quarter = warp & 3 <-> 32 queries of the tile, column half = warp >> 2)
// + 1 control warp; two CTAs per SM.
//   control warp: one elected lane streams 21.4 KB support "blobs" (pre-arranged UMMA operand images written by
//                 pack_supports_tc_kernel) L2 -> smem with 1-D bulk TMA through a 3-slot ring, issues GEMM1(j) and
//                 GEMM2(j-1) and signals completion with tcgen05.commit -> mbarrier;
//   query threads: FK -> A operand into shared memory; per chunk tcgen05.ld rho (96 columns), radial profile with packed
//                 FP32 (FFMA2/FMUL2) + MUFU.RCP, score in registers, coefficients split, packed to f16x2 and written back
//                 with tcgen05.st over the rho columns; epilogue tcgen05.ld G, J_FK^T, coalesced store.
#pragma once

#include <cuda_fp16.h>

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

namespace dc {

struct TcLayout {
  static constexpr int TM = 128;   // queries per tile == UMMA M
  static constexpr int NC = 96;    // support vectors per chunk == UMMA N of GEMM1
  static constexpr int FMAX = 14;  // features
  static constexpr int K1 = 48;    // GEMM1 K slots
  static constexpr int N2 = 32;    // GEMM2 N: [14 features, sum(cc), 0 | 14 feature corrections, 0, 0]
  static constexpr int ONES_ROW = 14;
  static constexpr int KS2 = NC / 8;  // GEMM2 instructions per chunk (8 supports x {ch, cl} each)
  // blob (bytes)
  static constexpr int B1_BYTES = NC * K1 * 2;        // [K1/8][NC][8] f16          9216
  static constexpr int B2_STEP_BYTES = N2 * 16 * 2;   // per 8 supports: [2][32][8]   1024
  static constexpr int OFF_B1 = 0;
  static constexpr int OFF_B2 = OFF_B1 + B1_BYTES;
  static constexpr int OFF_W = OFF_B2 + KS2 * B2_STEP_BYTES;  // 21504
  static constexpr int BLOB_BYTES = OFF_W + NC * 4;            // 21888 = 171 x 128
  static constexpr int TRAILER_FLOATS = 8;  // {max|s|^2, max|w| (bit patterns, atomicMax), Sx, 1/Sx, Sw, 1/Sw, 0, 0}
  static constexpr int RS = 3;              // ring slots
  // TMEM columns
  static constexpr int COL_STAGE = NC;      // per stage: rho, overwritten in place by the packed coefficients
  static constexpr int COL_G = 2 * COL_STAGE;  // 192 .. 223
  static constexpr int TMEM_COLS = 256;
  // shared memory (bytes)
  static constexpr int SM_BAR = 0;  // mbarriers
  static constexpr int SM_TMEM_SLOT = 128;
  static constexpr int SM_RING = 256;
  static constexpr int SM_A = SM_RING + RS * BLOB_BYTES;       // A [K1/8][128][8] f16
  static constexpr int QCAP = 128;                             // near-pair queue entries per warp
  static constexpr int QWARPS = 8;                             // query warps: 4 TMEM lane quarters x 2 column halves
  static constexpr int QTHREADS = QWARPS * 32;
  static constexpr int CTRL_WARP = QWARPS;
  static constexpr int THREADS = QTHREADS + 128;  // + the control warpgroup: warp 8 works, warps 9..11 only donate registers
  static constexpr int SM_XS = SM_A + TM * K1 * 2;             // features of the tile [128][16] f32 (near-pair path)
  // one region, three lives per tile: staged q [128][16] -> per-warp exact accumulators (feature gradient
  // [8][32][16] + score [8][32]) -> output records [128][17]
  static constexpr int SM_GEX = SM_XS + TM * 16 * 4;
  static constexpr int SM_QS = SM_GEX;
  static constexpr int GEX_BYTES = QWARPS * 32 * 17 * 4;        // 17408 >= TM * (DC_MAX_DOF + 1) * 4
  static constexpr int SM_ROWS = SM_GEX + GEX_BYTES;            // partial scores, thresholds [2][128]
  static constexpr int SM_QUEUE = SM_ROWS + 2 * TM * 4;
  static constexpr int SM_BYTES = SM_QUEUE + QWARPS * QCAP * 4;
};

struct TcArgs {
  dc_fk_desc fk;
  RadialConsts<float> rc;
  const unsigned char* blob;  // n_chunks x BLOB_BYTES + trailer
  const float* table;         // packed [-s | w] rows (dc_pack_supports), for the exact near-pair path
  const float* q;
  float* score;
  float* grad;
  const float* grad_out;
  long long* trace;  // optional clock64 timeline of CTA 0 (tools/probe/tc_probe.cu), 16 slots per chunk
  float* dbg;  // optional debug dump (tools/probe/tc_probe.cu): tile 0 rho [128][n_chunks*NC] then G [128][32]
  long long batch;
  long long score_ld;
  long long grad_ld;
  int n_sv;
  int n_feat;
  int n_in;
  int row_stride;  // of `table`
  int f_pad;
  int n_tiles;
  int n_chunks;
  float* bcast[DC_MAX_PEERS];  // n_bcast > 0: every fused record block is stored into each of these (row 0 = row 0 of the
  int n_bcast;                 // gathered buffer; `score` then points at THIS rank's block of the first one)
  float err_coef;  // delta(rho) <= err_coef * (|x|^2 + max|s|^2)
  float tol_pair;  // admissible |w|-relative error of one pair
};

// ---- tcgen05 / TMEM wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32 forms (tools/probe/umma_latency.cu)
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One lane of a converged warp (elect.sync): the form under which ptxas emits tcgen05.mma / TMA issue on the
// uniform datapath without a per-lane waterfall loop (235 -> 77 cycles per MMA, profiles/r01c_umma_issue_cost.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ bool mbar_test(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_test(addr, parity)) return;
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 60000000000LL) __trap();  // ~30 s: a protocol bug, not a slow launch
  }
}

// K-major, no-swizzle UMMA shared-memory descriptor: 8-row x 16-byte core matrices, `lbo` bytes between the two
// 16-byte K slabs of one instruction, `sbo` bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptors: D = F32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// 11-bit split of an fp32 value: hi keeps the top 11 significant bits (exact in f16 inside its normal range), lo = v - hi.
__device__ __forceinline__ float split_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
// {lo half = a, hi half = b} as f16x2
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float pow2_floor(float v) { return __uint_as_float(__float_as_uint(v) & 0x7f800000u); }

// ---- pack: support vectors -> chunk blobs ----------------------------------------------------------------------
// trailer[0], trailer[1] = max |s|^2, max |w| (as int bit patterns; zeroed by the caller before this kernel)
__global__ void __launch_bounds__(128) tc_scan_kernel(const float* __restrict__ s, const float* __restrict__ w, int n, int F,
                                                        int* __restrict__ trailer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ss = 0.f;
  for (int f = 0; f < F; ++f) ss = fmaf(s[(size_t)i * F + f], s[(size_t)i * F + f], ss);
  atomicMax(&trailer[0], __float_as_int(ss));
  atomicMax(&trailer[1], __float_as_int(fabsf(w[i])));
}

// Power-of-two scales: Sx^2 max|s|^2 <= 2^14 (so every f16 term of the supports is < 2^15), Sw max|w| <= 2^14.
__device__ __forceinline__ void tc_scales(const float* trailer, float& sx, float& sw) {
  const float ssmax = trailer[0], wmax = trailer[1];
  sx = (ssmax > 0.f) ? pow2_floor(sqrtf(16384.f / ssmax)) : 1.f;
  sw = (wmax > 0.f) ? pow2_floor(16384.f / wmax) : 1.f;
  sx = fminf(fmaxf(sx, 1.f / 1048576.f), 1048576.f);
  sw = fminf(fmaxf(sw, 1.f / 1048576.f), 1048576.f);
}

// One thread per (chunk, local support index).  s_feat[N, F] are the transformed supports, w[N] the weights.
__global__ void __launch_bounds__(128) pack_supports_tc_kernel(const float* __restrict__ s, const float* __restrict__ w,
                                                                 int n, int F, int n_chunks, unsigned char* __restrict__ blob) {
  using L = TcLayout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks * L::NC) return;
  float* trailer = reinterpret_cast<float*>(blob + (size_t)n_chunks * L::BLOB_BYTES);
  float sx, sw;
  tc_scales(trailer, sx, sw);
  if (idx == 0) {
    trailer[2] = sx;
    trailer[3] = 1.f / sx;
    trailer[4] = sw;
    trailer[5] = 1.f / sw;
    trailer[6] = 0.f;
    trailer[7] = 0.f;
  }
  const int j = idx / L::NC, r = idx - j * L::NC;
  unsigned char* b = blob + (size_t)j * L::BLOB_BYTES;
  float b1[L::K1], mainv[16], corrv[16];
#pragma unroll
  for (int k = 0; k < L::K1; ++k) b1[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) mainv[k] = corrv[k] = 0.f;
  float wv = 0.f;
  if (idx < n) {
    float ss = 0.f;
    for (int f = 0; f < F; ++f) {
      const float sv = sx * s[(size_t)idx * F + f];  // exact (power of two)
      const float m2 = -2.f * sv;
      const float hi = split_hi(m2);
      b1[f] = hi;                        // pairs with xh
      b1[16 + f] = hi * (1.f / 256.f);   // pairs with 256 xl
      b1[32 + f] = (m2 - hi) * 256.f;    // pairs with xh / 256
      const float sh = split_hi(sv);
      mainv[f] = sh;
      corrv[f] = sv - sh;
      ss = fmaf(sv, sv, ss);
    }
    const float s1 = split_hi(ss), s2 = split_hi(ss - s1), s3 = ss - s1 - s2;
    b1[14] = s1;  // x 1
    b1[15] = s2;  // x 1
    b1[30] = s3;  // x 1
    b1[31] = 1.f;  // x xx1
    b1[46] = 1.f;  // x xx2
    b1[47] = 1.f;  // x xx3
    mainv[L::ONES_ROW] = 1.f;
    wv = sw * w[idx];
  } else {
    b1[14] = 32768.f;  // padding rows: rho' >= 2^15 - |x'|^2, never near; weight 0 removes them from every sum
  }
  __half* b1p = reinterpret_cast<__half*>(b + L::OFF_B1);
#pragma unroll
  for (int k = 0; k < L::K1; ++k) b1p[(k >> 3) * (L::NC * 8) + r * 8 + (k & 7)] = __float2half_rn(b1[k]);
  // GEMM2 image of the 8-support step ks = r / 8: K slot i = r % 8 multiplies ch of support r, slot 8 + i its cl
  __half* b2p = reinterpret_cast<__half*>(b + L::OFF_B2 + (r >> 3) * L::B2_STEP_BYTES);
  const int i = r & 7;
#pragma unroll
  for (int f = 0; f < 16; ++f) {
    const __half hm = __float2half_rn(mainv[f]);
    b2p[0 * (L::N2 * 8) + f * 8 + i] = hm;                               // slot i,     row f:      sh (row 14: 1)
    b2p[1 * (L::N2 * 8) + f * 8 + i] = hm;                               // slot 8 + i, row f
    b2p[0 * (L::N2 * 8) + (16 + f) * 8 + i] = __float2half_rn(corrv[f]);  // slot i,     row 16 + f: s - sh
    b2p[1 * (L::N2 * 8) + (16 + f) * 8 + i] = __float2half_rn(0.f);
  }
  reinterpret_cast<float*>(b + L::OFF_W)[r] = wv;
}

#ifdef DC_TC_ENABLE_TRACE
#define DC_TC_TRACE(slot, gidx) \
  do { if (a.trace != nullptr && blockIdx.x == 1 && lane == 0 && (gidx) < 64) a.trace[(gidx) * 16 + (slot)] = clock64(); } while (0)
#else
#define DC_TC_TRACE(slot, gidx) do { } while (0)
#endif
#define DC_TC_TRACE_TILE(ev) do { if (warp == 0) DC_TC_TRACE((int)(t - t0), 50 + (ev)); } while (0)

// ---- the kernel -----------------------------------------------------------------------------------------------
enum TcMode { TC_SCORE = 0, TC_GRAD = 1 };

// Exact evaluation of queued near pairs: half-warp h takes one entry, lane f of the half takes feature f (direct
// difference against the fp32 support row), the 16 squares are summed with shuffles, and the pair's score / feature-
// gradient terms are added to the WARP'S OWN shared-memory accumulators of that row, strictly in queue (= column)
// order — so a row's result does not depend on its position in the batch nor on timing.  Four entries per half-warp
// are in flight at once so the L2 latency of the support rows overlaps.
__device__ __noinline__ void tc_drain_pairs(const TcArgs& a, const uint32_t* queue, int count, const float* xs,
                                            float* gacc_w, float* sacc_w, int lane) {
  const int h = lane >> 4, f = lane & 15;
  constexpr int U = 4;
  for (int head = 0; head < count; head += 2 * U) {
    bool valid[U];
    int row[U];
    float xv[U], tv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = head + 2 * u + h;
      valid[u] = idx < count;
      const uint32_t e = queue[valid[u] ? idx : head];
      row[u] = (int)(e >> 24);
      const int n = (int)(e & 0xffffffu);
      xv[u] = xs[row[u] * 16 + f];
      tv[u] = (f < a.row_stride) ? a.table[(size_t)n * a.row_stride + f] : 0.f;  // [-s | w] row
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float d = (f < a.n_feat) ? xv[u] + tv[u] : 0.f;
      float rho = d * d;
      rho += __shfl_xor_sync(0xffffffffu, rho, 8);
      rho += __shfl_xor_sync(0xffffffffu, rho, 4);
      rho += __shfl_xor_sync(0xffffffffu, rho, 2);
      rho += __shfl_xor_sync(0xffffffffu, rho, 1);
      const float wv = __shfl_sync(0xffffffffu, tv[u], a.f_pad, 16);
      float k, coef;
      radial_eval<KR_RQ2, float>(a.rc, rho, k, coef);
      const int r = row[u] & 31;
      const float cg = wv * coef * d, cs = wv * k;
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {  // entry 2u before entry 2u + 1
        if (h == ph && valid[u]) {
          if (f < a.n_feat) gacc_w[r * 16 + f] += cg;
          if (f == 15) sacc_w[r] += cs;
        }
        __syncwarp();
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(TcLayout::THREADS, 2) score_tc_kernel(const __grid_constant__ TcArgs a) {
  using L = TcLayout;
  constexpr int NC = L::NC, RS = L::RS, FM = L::FMAX, QT = L::QTHREADS;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::SM_BAR);
  uint64_t* bar_full = bars;             // [RS]  TMA -> MMA / query threads
  uint64_t* bar_free = bars + RS;        // [RS]  GEMM2 done -> TMA
  uint64_t* bar_rho = bars + 2 * RS;     // [2]   GEMM1 done -> query threads
  uint64_t* bar_cc = bars + 2 * RS + 2;  // [2]   query threads -> GEMM2
  uint64_t* bar_a = bars + 2 * RS + 4;   // [1]   A operand written -> GEMM1
  uint64_t* bar_g = bars + 2 * RS + 5;   // [1]   last GEMM2 of the tile done -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::SM_TMEM_SLOT);
  unsigned char* ring = smem + L::SM_RING;
  unsigned char* a_op = smem + L::SM_A;
  float* qs = reinterpret_cast<float*>(smem + L::SM_QS);
  float* os = qs;
  float* xs = reinterpret_cast<float*>(smem + L::SM_XS);       // [128][16] features of the tile (near-pair path)
  float* gacc = reinterpret_cast<float*>(smem + L::SM_GEX);    // [8][32][16] exact feature-gradient terms, per warp
  float* sacc = gacc + L::QWARPS * 32 * 16;                    // [8][32] exact score terms, per warp
  float* sc_p = reinterpret_cast<float*>(smem + L::SM_ROWS);   // [128] score partial of the second column half
  float* thr_s = sc_p + L::TM;                                 // [128] near threshold on rho'
  uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L::SM_QUEUE);

  const long long t0 = (long long)blockIdx.x * a.n_tiles / gridDim.x;
  const long long t1 = (long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x;
  const int nch = a.n_chunks;
  const uint32_t total = (uint32_t)((t1 - t0) * nch);  // < 2^31 (checked by launch_score_tc)

  if (warp == L::CTRL_WARP) {
    if (lane == 0) {
      for (int i = 0; i < RS; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_free[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_rho[i], 1);
        mbar_init(&bar_cc[i], QT);
      }
      mbar_init(bar_a, L::TM);
      mbar_init(bar_g, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, L::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Register reallocation between the warpgroups (setmaxnreg): the kernel is launched with 80 registers per thread
  // (12 warps x 2 CTAs per SM); the control warpgroup keeps 32 and the two query warpgroups grow to 104 (32 x 128 + 104 x 256 = the CTA's pool).
  if (warp >= L::CTRL_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    // ================= control warp: TMA producer + MMA issuer (one elected lane issues); warps 9..11 idle =========
    if (warp == L::CTRL_WARP && total > 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(NC);
      constexpr uint32_t idesc2 = umma_idesc_f16(L::N2);
      const uint32_t a_s = smem_u32(a_op);
      uint32_t issued = 0;
      auto issue = [&]() {
        const int slot = (int)(issued % RS);
        const int chunk = (int)(issued % nch);
        if (elect_one()) {
          mbar_expect_tx(&bar_full[slot], L::BLOB_BYTES);
          tma_bulk_g2s(ring + (size_t)slot * L::BLOB_BYTES, a.blob + (size_t)chunk * L::BLOB_BYTES, L::BLOB_BYTES,
                       &bar_full[slot]);
        }
        __syncwarp();
        ++issued;
      };
      for (int i = 0; i < RS && issued < total; ++i) issue();

      auto gemm2 = [&](uint32_t g, bool first_of_tile) {
        const int st = (int)(g & 1), slot = (int)(g % RS);
        mbar_wait_wd(&bar_cc[st], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        DC_TC_TRACE(3, g + 1);
        if (elect_one()) {
          if constexpr (MODE == TC_GRAD) {
            const uint32_t b_s = smem_u32(ring + (size_t)slot * L::BLOB_BYTES + L::OFF_B2);
            const uint32_t cc = tmem + st * L::COL_STAGE;
            const uint32_t d = tmem + L::COL_G;
#pragma unroll
            for (int ks = 0; ks < L::KS2; ++ks) {
              const uint64_t bd = umma_desc(b_s + ks * L::B2_STEP_BYTES, L::N2 * 16, 128);
              umma_f16_ts(d, cc + ks * 8, bd, idesc2, (first_of_tile && ks == 0) ? 0u : 1u);
            }
          }
          umma_commit(&bar_free[slot]);
        }
        __syncwarp();
      };

      uint32_t g = 0;
      for (long long t = t0; t < t1; ++t) {
        mbar_wait_wd(bar_a, (uint32_t)((t - t0) & 1));
        tc_fence_after();
        for (int j = 0; j < nch; ++j, ++g) {
          const int st = (int)(g & 1), slot = (int)(g % RS);
          mbar_wait_wd(&bar_full[slot], (uint32_t)((g / RS) & 1));
          tc_fence_after();
          DC_TC_TRACE(0, g);
          if (elect_one()) {
            const uint32_t b_s = smem_u32(ring + (size_t)slot * L::BLOB_BYTES + L::OFF_B1);
            const uint32_t d = tmem + st * L::COL_STAGE;
#pragma unroll
            for (int ks = 0; ks < L::K1 / 16; ++ks) {
              const uint64_t ad = umma_desc(a_s + ks * 2 * (L::TM * 16), L::TM * 16, 128);
              const uint64_t bd = umma_desc(b_s + ks * 2 * (NC * 16), NC * 16, 128);
              umma_f16_ss(d, ad, bd, idesc1, ks == 0 ? 0u : 1u);
            }
            umma_commit(&bar_rho[st]);
          }
          __syncwarp();
          DC_TC_TRACE(1, g);
          // refill the slot of chunk g-2 as soon as its GEMM2 (issued at the end of the previous iteration) has
          // completed: the copy of chunk g+1 then has the whole of chunk g-1's processing time to land
          while (issued < total && g >= 2 && issued - RS <= g - 2) {
            mbar_wait_wd(&bar_free[issued % RS], (uint32_t)(((issued / RS) + 1) & 1));
            issue();
          }
          DC_TC_TRACE(2, g);
          if (j >= 1) gemm2(g - 1, j == 1);
          DC_TC_TRACE(4, g);
        }
        gemm2(g - 1, nch == 1);
        if (elect_one()) umma_commit(bar_g);
        __syncwarp();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    // ================= query threads: row = 32 (warp & 3) + lane, column half = warp >> 2 ==========================
    const int row = ((warp & 3) << 5) | lane;
    const int hcol = warp >> 2;
    // owners run FK, the A operand and the epilogue of their row: the higher-numbered half, which the warp scheduler
    // favours — those phases are the CTA's critical path (the other half waits at the barriers)
    const bool owner = hcol == 1;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t* queue = queues + warp * L::QCAP;
    float* gacc_w = gacc + warp * 32 * 16;
    float* sacc_w = sacc + warp * 32;
    const int F = a.n_feat;
    const int n_out = 1 + (MODE == TC_GRAD ? a.n_in : 0);
    const bool fused = (MODE == TC_GRAD) ? (a.score_ld == a.grad_ld && a.score_ld == n_out && a.grad == a.score + 1)
                                         : (a.score_ld == 1);
    const float* trailer = reinterpret_cast<const float*>(a.blob + (size_t)nch * L::BLOB_BYTES);
    const float s2max = trailer[0], sx = trailer[2], inv_sx = trailer[3], inv_sw = trailer[5];
    const float c0s = a.rc.c0 * inv_sx * inv_sx;  // t = 1 + c0 rho = 1 + c0s rho'   (rho' = Sx^2 rho from GEMM1)
    uint32_t g = 0;
    for (long long t = t0; t < t1; ++t) {
      const long long b_base = t * L::TM;
      const int nq = (int)min((long long)L::TM, a.batch - b_base);
      // ---- stage the tile's configurations (coalesced), FK, A operand ----------------------------------------
      DC_TC_TRACE_TILE(0);
      {
        const float* src = a.q + (size_t)b_base * a.n_in;
        const int n_words = nq * a.n_in;
        if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4* s4 = reinterpret_cast<const float4*>(src);
          for (int i = tid; i < n_words / 4; i += QT) reinterpret_cast<float4*>(qs)[i] = s4[i];
        } else {
          for (int i = tid; i < n_words; i += QT) qs[i] = src[i];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float qv[DC_MAX_DOF];
      if (owner) {
#pragma unroll
        for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (row < nq && i < a.n_in) ? qs[row * a.n_in + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the staged configurations are consumed: the region becomes the
      {                                                 // exact accumulators of this tile
        float4* z = reinterpret_cast<float4*>(gacc_w + lane * 16);
        z[0] = z[1] = z[2] = z[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        sacc_w[lane] = 0.f;
      }
      if (owner) {
        float x[FM];
        {
          float xl[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) xl[i] = 0.f;
          if (row < nq) fk_forward<float>(a.fk, qv, xl, 1);
#pragma unroll
          for (int f = 0; f < FM; ++f) x[f] = (f < F) ? xl[f] : 0.f;
        }
        float xx = 0.f, xamax = 0.f;
#pragma unroll
        for (int f = 0; f < FM; ++f) {
          xx = fmaf(x[f], x[f], xx);
          xamax = fmaxf(xamax, fabsf(x[f]));
        }
        {
          float4* xr = reinterpret_cast<float4*>(xs + row * 16);
          xr[0] = make_float4(x[0], x[1], x[2], x[3]);
          xr[1] = make_float4(x[4], x[5], x[6], x[7]);
          xr[2] = make_float4(x[8], x[9], x[10], x[11]);
          xr[3] = make_float4(x[12], x[13], 0.f, 0.f);
        }
        // queries whose scaled features leave f16's range take the exact path for every pair (A row zeroed)
        const bool in_range = (sx * xamax < 16384.f) && (sx * sx * xx < 32768.f);
        {
          float v[L::K1];
          const float XX = in_range ? sx * sx * xx : 0.f;
#pragma unroll
          for (int k = 0; k < L::K1; ++k) v[k] = 0.f;
#pragma unroll
          for (int f = 0; f < FM; ++f) {
            const float xsc = in_range ? sx * x[f] : 0.f;
            const float hi = split_hi(xsc);
            v[f] = hi;                       // x (-2 sh)
            v[16 + f] = (xsc - hi) * 256.f;  // x (-2 sh) / 256
            v[32 + f] = hi * (1.f / 256.f);  // x 256 (-2 s)_lo
          }
          const float x1 = split_hi(XX), x2 = split_hi(XX - x1), x3 = XX - x1 - x2;
          v[14] = 1.f;  // x s1
          v[15] = 1.f;  // x s2
          v[30] = 1.f;  // x s3
          v[31] = x1;   // x 1
          v[46] = x2;   // x 1
          v[47] = x3;   // x 1
#pragma unroll
          for (int kc = 0; kc < L::K1 / 8; ++kc) {
            uint4 pk;
            pk.x = pack_f16x2(v[8 * kc + 0], v[8 * kc + 1]);
            pk.y = pack_f16x2(v[8 * kc + 2], v[8 * kc + 3]);
            pk.z = pack_f16x2(v[8 * kc + 4], v[8 * kc + 5]);
            pk.w = pack_f16x2(v[8 * kc + 6], v[8 * kc + 7]);
            reinterpret_cast<uint4*>(a_op)[kc * L::TM + row] = pk;
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_a);
        // near-pair threshold of this query, on rho' (file header): rho' below it is recomputed exactly
        const float drho = a.err_coef * (xx + s2max);
        const float tcrit = cbrtf(fmaxf(-a.rc.grad_scale * 0.5f * drho / a.tol_pair, 1.f));  // (gamma drho / tol)^(1/3)
        thr_s[row] = in_range ? sx * sx * (tcrit - 1.f) / a.rc.c0 : 3.0e38f;
      }
      DC_TC_TRACE_TILE(1);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // xs, thr_s visible; qs free for the output records
      DC_TC_TRACE_TILE(2);
      const float thr = thr_s[row];

      P2 sc2(0.f, 0.f);
      int qcount = 0;
      for (int j = 0; j < nch; ++j, ++g) {
        const int st = (int)(g & 1), slot = (int)(g % RS);
        const float* wsm = reinterpret_cast<const float*>(ring + (size_t)slot * L::BLOB_BYTES + L::OFF_W) + hcol * (NC / 2);
        if (warp == 0) DC_TC_TRACE(8, g);
        mbar_wait_wd(&bar_full[slot], (uint32_t)((g / RS) & 1));
        if (warp == 0) DC_TC_TRACE(9, g);
        mbar_wait_wd(&bar_rho[st], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (warp == 0) DC_TC_TRACE(10, g);
        const uint32_t tcol = tm_lane + st * L::COL_STAGE + hcol * (NC / 2);
        uint32_t r[NC / 2];
        tmem_ld16(tcol, r);
        tmem_ld16(tcol + 16, r + 16);
        tmem_ld16(tcol + 32, r + 32);
#pragma unroll
        for (int bt = 0; bt < 3; ++bt) {
          if (bt == 0) {
            tmem_wait_ld();
            if (warp == 0) DC_TC_TRACE(11, g);
          }
          uint32_t* rb = r + bt * 16;
          const int col0 = bt * 16;
          if (a.dbg != nullptr && t == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              a.dbg[(size_t)row * (nch * NC) + j * NC + hcol * (NC / 2) + col0 + c] = __uint_as_float(rb[c]) * inv_sx * inv_sx;
          }
          // ---- near pairs: queued for exact evaluation, removed from the tensor-core sums (rho' := huge -> u = 0) ----
          float mn = fminf(__uint_as_float(rb[0]), __uint_as_float(rb[1]));
#pragma unroll
          for (int c = 2; c < 16; c += 2) mn = fminf(mn, fminf(__uint_as_float(rb[c]), __uint_as_float(rb[c + 1])));
          if (__any_sync(0xffffffffu, mn < thr)) {
            uint32_t nearmask = 0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const bool near = __uint_as_float(rb[c]) < thr;
              nearmask |= near ? (1u << c) : 0u;
              rb[c] = near ? 0x5d000000u : rb[c];  // 5.8e17 -> u ~ 1e-18 (and t0 t1 stays finite): the pair leaves the sums
            }
            const int n0 = j * NC + hcol * (NC / 2) + col0;
            uint32_t todo = __reduce_or_sync(0xffffffffu, nearmask);  // columns some lane of the warp flagged
#pragma unroll 1
            while (todo != 0) {
              const int c = __ffs((int)todo) - 1;
              todo &= todo - 1;
              if (n0 + c >= a.n_sv) break;  // padding columns (only reachable for out-of-range queries): zero weight
              const bool near = (nearmask >> c) & 1u;
              const uint32_t bal = __ballot_sync(0xffffffffu, near);
              if (near) queue[qcount + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)row << 24) | (uint32_t)(n0 + c);
              qcount += __popc(bal);
              if (qcount > L::QCAP - 32) {
                __syncwarp();
                tc_drain_pairs(a, queue, qcount, xs, gacc_w, sacc_w, lane);
                __syncwarp();
                qcount = 0;
              }
            }
          }
          // ---- all pairs: radial profile on the tensor-core rho', packed over column pairs ---------------------
          uint32_t outp[16];  // per 8 supports: 4 x f16x2 ch, 4 x f16x2 cl  (GEMM2 K slots 0..7, 8..15)
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(wsm + col0 + c);
            const P2 rho2(__uint_as_float(rb[c]), __uint_as_float(rb[c + 1]));
            const P2 tt = pfma_bb(rho2, c0s, 1.0f);
            const float rr = fast_rcp(tt.lo() * tt.hi());  // one MUFU per column pair: 1/t0 = t1 / (t0 t1)
            const P2 u = pmul_b(P2(tt.hi(), tt.lo()), rr);
            const P2 k = pmul(u, u);
            const P2 wk = pmul(P2(w2), k);
            sc2 = padd(sc2, wk);
            if constexpr (MODE == TC_GRAD) {
              const P2 cc = pmul(wk, u);
              const P2 ch(split_hi(cc.lo()), split_hi(cc.hi()));
              const P2 cl = padd(cc, P2(-ch.lo(), -ch.hi()));
              const int o = (c >> 3) * 8 + ((c & 7) >> 1);
              outp[o] = pack_f16x2(ch.lo(), ch.hi());
              outp[o + 4] = pack_f16x2(cl.lo(), cl.hi());
            }
          }
          if constexpr (MODE == TC_GRAD) tmem_st16(tcol + col0, outp);
        }
        if (warp == 0) DC_TC_TRACE(12, g);
        if constexpr (MODE == TC_GRAD) tmem_wait_st();
        tc_fence_before();
        if (warp == 0) DC_TC_TRACE(13, g);
        mbar_arrive(&bar_cc[st]);
      }
      DC_TC_TRACE_TILE(3);
      if (qcount > 0) {
        __syncwarp();
        tc_drain_pairs(a, queue, qcount, xs, gacc_w, sacc_w, lane);
      }
      DC_TC_TRACE_TILE(4);
      if (!owner) sc_p[row] = sc2.lo() + sc2.hi();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // exact terms and second-half partial scores are complete
      DC_TC_TRACE_TILE(5);

      // ---- epilogue (row owners): G from TMEM, feature gradient, J_FK^T, records into shared memory -----------------
      if (owner) {
        float gx[DC_MAX_DOF], xl[DC_MAX_DOF];
#pragma unroll
        for (int i = 0; i < DC_MAX_DOF; ++i) {
          gx[i] = 0.f;
          xl[i] = xs[row * 16 + i];
        }
        mbar_wait_wd(bar_g, (uint32_t)((t - t0) & 1));
        tc_fence_after();
        if constexpr (MODE == TC_GRAD) {
          uint32_t gm[16], gc[16];
          tmem_ld16(tm_lane + L::COL_G, gm);
          tmem_ld16(tm_lane + L::COL_G + 16, gc);
          tmem_wait_ld();
          if (a.dbg != nullptr && t == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              a.dbg[(size_t)L::TM * (nch * NC) + row * 32 + c] = __uint_as_float(gm[c]);
              a.dbg[(size_t)L::TM * (nch * NC) + row * 32 + 16 + c] = __uint_as_float(gc[c]);
              a.dbg[(size_t)L::TM * (nch * NC) + L::TM * 32 + row * 16 + c] = xl[c];
            }
          }
          const float csum = __uint_as_float(gm[L::ONES_ROW]);
#pragma unroll
          for (int f = 0; f < FM; ++f) {
            const float gsum = (__uint_as_float(gm[f]) + __uint_as_float(gc[f])) * inv_sx;  // sum cc' s
            const float gex = gacc_w[lane * 16 + f - 4 * 32 * 16] + gacc_w[lane * 16 + f];  // column halves 0 + 1
            gx[f] = a.rc.grad_scale * (fmaf(xl[f], csum, -gsum) * inv_sw + gex);
          }
        }
        tc_fence_before();
        const float score =
            a.rc.score_scale * ((sc_p[row] + (sc2.lo() + sc2.hi())) * inv_sw + (sacc_w[lane - 4 * 32] + sacc_w[lane]));
        asm volatile("bar.sync 2, 128;" ::: "memory");  // every owner has read its accumulators: the region becomes `os`
        if (row < nq) {
          float* rec = os + row * n_out;
          rec[0] = score;
          if constexpr (MODE == TC_GRAD) {
            const float scale = (a.grad_out != nullptr) ? a.grad_out[b_base + row] : 1.f;
            if (a.fk.type == DC_FK_NONE) {
              for (int f = 0; f < F; ++f) rec[1 + f] = scale * gx[f];
            } else {
              float gq[DC_MAX_DOF];
#pragma unroll
              for (int i = 0; i < DC_MAX_DOF; ++i) gq[i] = 0.f;
              fk_vjp<float>(a.fk, qv, xl, 1, gx, 1, gq);
              for (int i = 0; i < a.n_in; ++i) rec[1 + i] = scale * gq[i];
            }
          }
        }
      } else {
        // the accumulator must not be overwritten by the next tile's first GEMM2 before the owners have read it: the
        // owners' wait on bar_g above orders that (their next arrival on bar_cc comes after this epilogue)
      }
      DC_TC_TRACE_TILE(6);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      DC_TC_TRACE_TILE(7);
      if (fused) {
        // n_bcast > 0: the same block goes to every rank's gathered buffer (peer stores over NVLink) — the all-gather of
        // the multi-GPU path, overlapped with the other tiles' arithmetic
        const int n_dst = a.n_bcast > 0 ? a.n_bcast : 1;
        const size_t off = (size_t)(a.score - (a.n_bcast > 0 ? a.bcast[0] : a.score)) + (size_t)b_base * n_out;
        const int n_words = nq * n_out;
        if (a.n_bcast > 0 && (n_words & 3) == 0) {
          // one bulk TMA store of the whole block per destination (shared -> peer global, large NVLink packets, no
          // thread is held by the transfer); the block is released below once the copies have read it
          if (tid == 0) {
            fence_proxy_async();
            for (int k = 0; k < n_dst; ++k)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.bcast[k] + off),
                           "r"(smem_u32(os)), "r"(n_words * 4)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
        } else {
          for (int k = 0; k < n_dst; ++k) {
            float* dst = (a.n_bcast > 0 ? a.bcast[k] : a.score) + off;
            if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              for (int i = tid; i < n_words / 4; i += QT) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(os)[i];
            } else {
              for (int i = tid; i < n_words; i += QT) dst[i] = os[i];
            }
          }
        }
      } else {
        if (tid < nq) a.score[(size_t)(b_base + tid) * a.score_ld] = os[tid * n_out];
        if constexpr (MODE == TC_GRAD) {
          for (int i = tid; i < nq * a.n_in; i += QT) {
            const int tq = i / a.n_in, c = i - tq * a.n_in;
            a.grad[(size_t)(b_base + tq) * a.grad_ld + c] = os[tq * n_out + 1 + c];
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // qs / os are rewritten by the next tile
      DC_TC_TRACE_TILE(8);
    }
  }

  if (tid == 0 && a.n_bcast > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // peer stores performed
  tc_fence_before();
  __syncthreads();
  if (warp == L::CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

inline int tc_n_chunks(long long n_sv) { return (int)((n_sv + TcLayout::NC - 1) / TcLayout::NC); }
inline size_t tc_blob_bytes(long long n_sv) {
  return (size_t)tc_n_chunks(n_sv) * TcLayout::BLOB_BYTES + TcLayout::TRAILER_FLOATS * 4;
}

// Packs S_feat[N,F], w[N] into `blob` (tc_blob_bytes(N) bytes, 128-byte aligned).  Three launches on `stream`.
inline int launch_pack_supports_tc(const float* s_feat, const float* w, long long n, int F, unsigned char* blob,
                                   cudaStream_t stream) {
  using L = TcLayout;
  const int nch = tc_n_chunks(n);
  int* trailer = reinterpret_cast<int*>(blob + (size_t)nch * L::BLOB_BYTES);
  DC_CUDA_OK(cudaMemsetAsync(trailer, 0, L::TRAILER_FLOATS * 4, stream));
  tc_scan_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(s_feat, w, (int)n, F, trailer);
  DC_LAUNCH_CHECK();
  pack_supports_tc_kernel<<<(unsigned)((nch * L::NC + 127) / 128), 128, 0, stream>>>(s_feat, w, (int)n, F, nch, blob);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

template <int MODE>
int launch_score_tc(TcArgs& a, int num_sms, cudaStream_t stream) {
  using L = TcLayout;
  a.n_tiles = (int)ceil_div64(a.batch, L::TM);
  a.n_chunks = tc_n_chunks(a.n_sv);
  const int grid = (int)min((long long)2 * num_sms, (long long)a.n_tiles);
  if (((long long)a.n_tiles / grid + 1) * a.n_chunks >= (1LL << 31) || a.n_sv >= (1 << 24)) return DC_ERR_UNSUPPORTED;
  auto kern = score_tc_kernel<MODE>;
  DC_SET_FUNC_ATTR_ONCE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SM_BYTES);
  kern<<<grid, L::THREADS, L::SM_BYTES, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc
