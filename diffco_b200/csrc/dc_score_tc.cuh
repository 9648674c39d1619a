// Fused collision-score kernel, tensor-core form (tcgen05 + TMEM + bulk TMA; RQKernel(p = 2), one class, fp32, F <= 14).
//
//   score[b] = sum_n w_n k(rho_bn)        g_x[b] = -2 gamma sum_n w_n u^3 (x_b - s_n)        rho_bn = |x_b - s_n|^2
//
// Same contract as score_tq_kernel (dc_score_tq.cuh) for diffco/kernel_perceptrons.py:362-370 (DiffCo.score with
// diffco/kernel.py:17-29 and a diffco/model.py feature map) and its autograd backward, but the two contractions
// run on the 5th-generation tensor cores instead of the FP32 pipe:
//
//   GEMM1  rho[128 x NC]  = A[128 x 16] . B1[NC x 16]^T      A  = [x_0..x_13 | |x|^2 | 1]    B1 = [-2 s | 1 | |s|^2]
//   GEMM2  G  [128 x 16] += CC[128 x NC] . B2[16 x NC]^T     CC = w u^3 (from TMEM)          B2 = [s | 1 | 0]
//
// so that g_x = -2 gamma (x G[:,14] - G[:,0..13]).  Both are kind::tf32 with every operand split into two TF32 terms
// (hi + lo, three passes hi.hi + lo.hi + hi.lo: products carry ~22 bits, accumulation is fp32 in TMEM).  The expansion
// |x|^2 + |s|^2 - 2 x.s carries an ABSOLUTE error of ~1e-5 in rho, which matters only where k'(rho) is large, i.e. for
// the few pairs with small rho: those (rho below a per-query threshold derived from the error bound) are recomputed
// exactly with direct differences on the FP32 pipe and removed from the tensor-core gradient.  For every other pair
// the error contribution is below 2e-7 of a unit weight.
//
// CTA = 128 query threads (thread i <-> query i of the tile <-> TMEM lane i) + 1 control warp; two CTAs per SM.
//   control lane: streams 12.5 KB support "blobs" (pre-arranged UMMA operand images, dc_pack_supports_tc) L2 -> smem
//                 with 1-D bulk TMA through a 4-slot ring, issues GEMM1(j) and GEMM2(j-1) with tcgen05.mma, and
//                 signals completion with tcgen05.commit -> mbarrier;
//   query threads: FK -> A operand (hi/lo) into shared memory; per chunk tcgen05.ld rho (48 columns), radial profile
//                 with packed FP32 (FFMA2/FMUL2) + MUFU.RCP, score in registers, coefficients split hi/lo and
//                 written back with tcgen05.st over the rho columns; epilogue tcgen05.ld G, J_FK^T, coalesced store.
#pragma once

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

namespace dc {

struct TcLayout {
  static constexpr int TM = 128;              // queries per tile == UMMA M
  static constexpr int NC = 48;               // support vectors per chunk == UMMA N of GEMM1
  static constexpr int FMAX = 14;             // features (K = 16 = F + |x|^2 + 1)
  static constexpr int K1 = 16;
  static constexpr int N2 = 16;               // GEMM2 N: 14 feature columns + sum(cc) + 0
  static constexpr int ONES_ROW = 14;
  // blob (floats)
  static constexpr int B1_FLOATS = NC * K1;   // 768: [k/4][NC][4]
  static constexpr int B2_FLOATS = N2 * NC;   // 768: [n/4][16][4]
  static constexpr int OFF_B1HI = 0;
  static constexpr int OFF_B1LO = OFF_B1HI + B1_FLOATS;
  static constexpr int OFF_B2HI = OFF_B1LO + B1_FLOATS;
  static constexpr int OFF_B2LO = OFF_B2HI + B2_FLOATS;
  static constexpr int OFF_W = OFF_B2LO + B2_FLOATS;
  static constexpr int BLOB_FLOATS = 3136;    // 12544 B = 98 x 128 B (OFF_W + NC = 3120, padded)
  static constexpr int BLOB_BYTES = BLOB_FLOATS * 4;
  static constexpr int RS = 4;                // ring slots
  // TMEM columns
  static constexpr int COL_STAGE = 2 * NC;    // per stage: rho / cc_hi [0,NC), cc_lo [NC, 2NC)
  static constexpr int COL_G = 2 * COL_STAGE; // 192
  static constexpr int TMEM_COLS = 256;
  // shared memory (bytes)
  static constexpr int SM_BAR = 0;            // 16 mbarriers
  static constexpr int SM_TMEM_SLOT = 128;
  static constexpr int SM_RING = 256;
  static constexpr int SM_A = SM_RING + RS * BLOB_BYTES;          // A hi [4][128][4], A lo
  static constexpr int SM_QS = SM_A + 2 * TM * K1 * 4;             // staged q [128][16]
  static constexpr int SM_OS = SM_QS + TM * DC_MAX_DOF * 4;        // output records [128][17]
  static constexpr int SM_BYTES = SM_OS + TM * (DC_MAX_DOF + 1) * 4;
};

struct TcArgs {
  dc_fk_desc fk;
  RadialConsts<float> rc;
  const float* blob;     // n_chunks x BLOB_FLOATS + trailer {max |s|^2}
  const float* table;    // packed [-s | w] rows (dc_pack_supports), for the exact near-pair path
  const float* q;
  float* score;
  float* grad;
  const float* grad_out;
  float* dbg;            // optional debug dump (tools/tc_probe.cu): tile 0 rho [128][n_chunks*NC] then G [128][16]
  long long batch;
  long long score_ld;
  long long grad_ld;
  int n_sv;
  int n_feat;
  int n_in;
  int row_stride;        // of `table`
  int f_pad;
  int n_tiles;
  int n_chunks;
  float err_coef;        // delta(rho) <= err_coef * (|x|^2 + max|s|^2)
  float tol_pair;        // admissible |w|-relative error of one pair
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
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One lane of a converged warp (elect.sync): the form under which ptxas emits tcgen05.mma / TMA issue on the
// uniform datapath without a per-lane waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// K-major, no-swizzle UMMA shared-memory descriptor: 8-row x 16-byte core matrices, `lbo` bytes between the two
// 16-byte K slabs of one instruction, `sbo` bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- pack: support vectors -> chunk blobs ----------------------------------------------------------------------
// One thread per (chunk, local support index).  s_feat[N, F] are the transformed supports, w[N] the weights.
__global__ void __launch_bounds__(128) pack_supports_tc_kernel(const float* __restrict__ s, const float* __restrict__ w,
                                                                 int n, int F, int n_chunks, float* __restrict__ blob) {
  using L = TcLayout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks * L::NC) return;
  const int j = idx / L::NC, r = idx - j * L::NC;
  float* b = blob + (size_t)j * L::BLOB_FLOATS;
  float v1[L::K1], v2[L::N2];
  float wv = 0.f;
#pragma unroll
  for (int k = 0; k < L::K1; ++k) v1[k] = 0.f;
#pragma unroll
  for (int k = 0; k < L::N2; ++k) v2[k] = 0.f;
  if (idx < n) {
    float ss = 0.f;
    for (int f = 0; f < F; ++f) {
      const float sv = s[(size_t)idx * F + f];
      v1[f] = -2.f * sv;
      v2[f] = sv;
      ss = fmaf(sv, sv, ss);
    }
    v1[14] = 1.f;
    v1[15] = ss;
    v2[L::ONES_ROW] = 1.f;
    wv = w[idx];
    atomicMax(reinterpret_cast<int*>(blob + (size_t)n_chunks * L::BLOB_FLOATS), __float_as_int(ss));
  } else {
    v1[15] = 1e30f;  // padding rows: rho = 1e30 -> u = 0; weight 0
  }
#pragma unroll
  for (int k = 0; k < L::K1; ++k) {
    const float hi = tf32_rn(v1[k]);
    const int o = (k >> 2) * (L::NC * 4) + r * 4 + (k & 3);
    b[L::OFF_B1HI + o] = hi;
    b[L::OFF_B1LO + o] = v1[k] - hi;
  }
#pragma unroll
  for (int f = 0; f < L::N2; ++f) {
    const float hi = tf32_rn(v2[f]);
    const int o = (r >> 2) * (L::N2 * 4) + f * 4 + (r & 3);
    b[L::OFF_B2HI + o] = hi;
    b[L::OFF_B2LO + o] = v2[f] - hi;
  }
  b[L::OFF_W + r] = wv;
  if (r < L::BLOB_FLOATS - L::OFF_W - L::NC) b[L::OFF_W + L::NC + r] = 0.f;
}

// ---- the kernel -----------------------------------------------------------------------------------------------
enum TcMode { TC_SCORE = 0, TC_GRAD = 1 };

template <int MODE>
__global__ void __launch_bounds__(160, 2) score_tc_kernel(const __grid_constant__ TcArgs a) {
  using L = TcLayout;
  constexpr int NC = L::NC, RS = L::RS, FM = L::FMAX;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::SM_BAR);
  uint64_t* bar_full = bars;              // [RS]  TMA -> MMA / query threads
  uint64_t* bar_free = bars + RS;         // [RS]  GEMM2 done -> TMA
  uint64_t* bar_rho = bars + 2 * RS;      // [2]   GEMM1 done -> query threads
  uint64_t* bar_cc = bars + 2 * RS + 2;   // [2]   query threads -> GEMM2
  uint64_t* bar_a = bars + 2 * RS + 4;    // [1]   A operand written -> GEMM1
  uint64_t* bar_g = bars + 2 * RS + 5;    // [1]   last GEMM2 of the tile done -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::SM_TMEM_SLOT);
  float* ring = reinterpret_cast<float*>(smem + L::SM_RING);
  float* a_hi = reinterpret_cast<float*>(smem + L::SM_A);
  float* a_lo = a_hi + L::TM * L::K1;
  float* qs = reinterpret_cast<float*>(smem + L::SM_QS);
  float* os = reinterpret_cast<float*>(smem + L::SM_OS);

  const long long t0 = (long long)blockIdx.x * a.n_tiles / gridDim.x;
  const long long t1 = (long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x;
  const int nch = a.n_chunks;
  const long long total = (t1 - t0) * nch;

  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < RS; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_free[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_rho[i], 1);
        mbar_init(&bar_cc[i], L::TM);
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

  if (warp == 4) {
    // ================= control lane: TMA producer + MMA issuer =================================================
    if (total > 0) {
      constexpr uint32_t idesc1 = umma_idesc_tf32(NC);
      constexpr uint32_t idesc2 = umma_idesc_tf32(L::N2);
      const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo);
      long long issued = 0;
      auto issue = [&]() {
        const int slot = (int)(issued % RS);
        const int chunk = (int)(issued % nch);
        if (elect_one()) {
          mbar_expect_tx(&bar_full[slot], L::BLOB_BYTES);
          tma_bulk_g2s(ring + (size_t)slot * L::BLOB_FLOATS, a.blob + (size_t)chunk * L::BLOB_FLOATS, L::BLOB_BYTES,
                       &bar_full[slot]);
        }
        __syncwarp();
        ++issued;
      };
      for (int i = 0; i < RS && issued < total; ++i) issue();

      auto gemm2 = [&](long long g, bool first_of_tile) {
        const int st = (int)(g & 1), slot = (int)(g % RS);
        mbar_wait_wd(&bar_cc[st], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
         if constexpr (MODE == TC_GRAD) {
          const uint32_t b_s = smem_u32(ring + (size_t)slot * L::BLOB_FLOATS);
          const uint32_t cc_hi = tmem + st * L::COL_STAGE, cc_lo = cc_hi + NC;
          const uint32_t d = tmem + L::COL_G;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t at = (pass == 1) ? cc_lo : cc_hi;
            const uint32_t bb = b_s + 4 * ((pass == 2) ? L::OFF_B2LO : L::OFF_B2HI);
#pragma unroll
            for (int ks = 0; ks < NC / 8; ++ks) {
              const uint64_t bd = umma_desc(bb + ks * 2 * (L::N2 * 16), L::N2 * 16, 128);
              umma_ts(d, at + ks * 8, bd, idesc2, (first_of_tile && pass == 0 && ks == 0) ? 0u : 1u);
            }
          }
         }
         umma_commit(&bar_free[slot]);
        }
        __syncwarp();
      };

      long long g = 0;
      for (long long t = t0; t < t1; ++t) {
        mbar_wait_wd(bar_a, (uint32_t)((t - t0) & 1));
        tc_fence_after();
        for (int j = 0; j < nch; ++j, ++g) {
          const int st = (int)(g & 1), slot = (int)(g % RS);
          mbar_wait_wd(&bar_full[slot], (uint32_t)((g / RS) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_s = smem_u32(ring + (size_t)slot * L::BLOB_FLOATS);
            const uint32_t d = tmem + st * L::COL_STAGE;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t aa = (pass == 1) ? a_lo_s : a_hi_s;
              const uint32_t bb = b_s + 4 * ((pass == 2) ? L::OFF_B1LO : L::OFF_B1HI);
#pragma unroll
              for (int ks = 0; ks < L::K1 / 8; ++ks) {
                const uint64_t ad = umma_desc(aa + ks * 2 * (L::TM * 16), L::TM * 16, 128);
                const uint64_t bd = umma_desc(bb + ks * 2 * (NC * 16), NC * 16, 128);
                umma_ss(d, ad, bd, idesc1, (pass == 0 && ks == 0) ? 0u : 1u);
              }
            }
            umma_commit(&bar_rho[st]);
          }
          __syncwarp();
          if (j >= 1) gemm2(g - 1, j == 1);
          // refill the slot of chunk g-2 (its GEMM2 was issued one iteration ago and has long completed)
          while (issued < total && issued - RS <= g - 2) {
            mbar_wait_wd(&bar_free[issued % RS], (uint32_t)(((issued / RS) + 1) & 1));
            issue();
          }
        }
        gemm2(g - 1, nch == 1);
        if (elect_one()) umma_commit(bar_g);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================= query threads ===========================================================================
    const uint32_t tm_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const int F = a.n_feat;
    const int n_out = 1 + (MODE == TC_GRAD ? a.n_in : 0);
    const bool fused = (MODE == TC_GRAD) ? (a.score_ld == a.grad_ld && a.score_ld == n_out && a.grad == a.score + 1)
                                         : (a.score_ld == 1);
    const float s2max = a.blob[(size_t)nch * L::BLOB_FLOATS];
    long long g = 0;
    for (long long t = t0; t < t1; ++t) {
      const long long b_base = t * L::TM;
      const int nq = (int)min((long long)L::TM, a.batch - b_base);
      // ---- stage the tile's configurations (coalesced), FK, A operand ----------------------------------------
      {
        const float* src = a.q + (size_t)b_base * a.n_in;
        const int n_words = nq * a.n_in;
        if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4* s4 = reinterpret_cast<const float4*>(src);
          for (int i = tid; i < n_words / 4; i += L::TM) reinterpret_cast<float4*>(qs)[i] = s4[i];
        } else {
          for (int i = tid; i < n_words; i += L::TM) qs[i] = src[i];
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float qv[DC_MAX_DOF], x[FM];
      {
        float xl[DC_MAX_DOF];
#pragma unroll
        for (int i = 0; i < DC_MAX_DOF; ++i) {
          qv[i] = (tid < nq && i < a.n_in) ? qs[tid * a.n_in + i] : 0.f;
          xl[i] = 0.f;
        }
        if (tid < nq) fk_forward<float>(a.fk, qv, xl, 1);
#pragma unroll
        for (int f = 0; f < FM; ++f) x[f] = (f < F) ? xl[f] : 0.f;
      }
      float xx = 0.f;
#pragma unroll
      for (int f = 0; f < FM; ++f) xx = fmaf(x[f], x[f], xx);
      {
        float v[L::K1];
#pragma unroll
        for (int f = 0; f < FM; ++f) v[f] = x[f];
        v[14] = xx;
        v[15] = 1.f;
#pragma unroll
        for (int kc = 0; kc < L::K1 / 4; ++kc) {
          float4 h, l;
          h.x = tf32_rn(v[4 * kc]);
          h.y = tf32_rn(v[4 * kc + 1]);
          h.z = tf32_rn(v[4 * kc + 2]);
          h.w = tf32_rn(v[4 * kc + 3]);
          l.x = v[4 * kc] - h.x;
          l.y = v[4 * kc + 1] - h.y;
          l.z = v[4 * kc + 2] - h.z;
          l.w = v[4 * kc + 3] - h.w;
          reinterpret_cast<float4*>(a_hi)[kc * L::TM + tid] = h;
          reinterpret_cast<float4*>(a_lo)[kc * L::TM + tid] = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_a);

      // near-pair threshold of this query: rho below it is recomputed exactly (file header)
      const float drho = a.err_coef * (xx + s2max);
      const float tcrit = cbrtf(fmaxf(-a.rc.grad_scale * 0.5f * drho / a.tol_pair, 1.f));  // (gamma drho / tol)^(1/3)
      const float thr = (tcrit - 1.f) / a.rc.c0;

      P2 sc2(0.f, 0.f);
      float sc_ex = 0.f;
      float gex[FM];
#pragma unroll
      for (int f = 0; f < FM; ++f) gex[f] = 0.f;

      for (int j = 0; j < nch; ++j, ++g) {
        const int st = (int)(g & 1), slot = (int)(g % RS);
        const float* wsm = ring + (size_t)slot * L::BLOB_FLOATS + L::OFF_W;
        mbar_wait_wd(&bar_full[slot], (uint32_t)((g / RS) & 1));
        mbar_wait_wd(&bar_rho[st], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        const uint32_t tcol = tm_lane + st * L::COL_STAGE;
        uint32_t r[NC];
        tmem_ld16(tcol, r);
        tmem_ld16(tcol + 16, r + 16);
        tmem_ld16(tcol + 32, r + 32);
        tmem_wait_ld();
        if (a.dbg != nullptr && t == 0) {
#pragma unroll
          for (int c = 0; c < NC; ++c) a.dbg[(size_t)tid * (nch * NC) + j * NC + c] = __uint_as_float(r[c]);
        }
#pragma unroll
        for (int bt = 0; bt < NC / 16; ++bt) {
          uint32_t* rb = r + bt * 16;
          // ---- near pairs: exact direct-difference evaluation on the FP32 pipe ---------------------------------
          float mn = __uint_as_float(rb[0]);
#pragma unroll
          for (int c = 1; c < 16; ++c) mn = fminf(mn, __uint_as_float(rb[c]));
          uint32_t nearmask = 0;
          if (__any_sync(0xffffffffu, mn < thr)) {
#pragma unroll
            for (int c = 0; c < 16; ++c) nearmask |= (__uint_as_float(rb[c]) < thr) ? (1u << c) : 0u;
            uint32_t todo = __reduce_or_sync(0xffffffffu, nearmask);  // columns some lane of the warp needs
            while (todo != 0) {
              const int c = __ffs((int)todo) - 1;
              todo &= todo - 1;
              const int n = j * NC + bt * 16 + c;  // < n_sv: padding rows carry rho = 1e30
              const float* row = a.table + (size_t)n * a.row_stride;
              float d[FM], rho = 0.f;
#pragma unroll
              for (int f = 0; f < FM; ++f) {
                d[f] = (f < F) ? x[f] + row[f] : 0.f;
                rho = fmaf(d[f], d[f], rho);
              }
              float k, coef;
              radial_eval<KR_RQ2, float>(a.rc, rho, k, coef);
              if ((nearmask >> c) & 1u) {
                const float wv = row[a.f_pad];
                sc_ex = fmaf(wv, k, sc_ex);
                const float cw = wv * coef;
#pragma unroll
                for (int f = 0; f < FM; ++f) gex[f] = fmaf(cw, d[f], gex[f]);
              }
            }
          }
          // ---- all pairs: radial profile on the tensor-core rho, packed over column pairs -----------------------
          uint32_t hi[16], lo[16];
          P2 wk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(wsm + bt * 16 + c);
            const P2 rho2(__uint_as_float(rb[c]), __uint_as_float(rb[c + 1]));
            const P2 tt = pfma_bb(rho2, a.rc.c0, 1.0f);
            const P2 u(fast_rcp(tt.lo()), fast_rcp(tt.hi()));
            const P2 k = pmul(u, u);
            wk[c / 2] = pmul(P2(w2), k);
            if constexpr (MODE == TC_GRAD) {
              const P2 cc = pmul(wk[c / 2], u);
              hi[c] = __float_as_uint(cc.lo()) & 0xffffe000u;
              hi[c + 1] = __float_as_uint(cc.hi()) & 0xffffe000u;
              const P2 l = padd(cc, P2(-__uint_as_float(hi[c]), -__uint_as_float(hi[c + 1])));
              lo[c] = __float_as_uint(l.lo());
              lo[c + 1] = __float_as_uint(l.hi());
            }
          }
          if (nearmask != 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if ((nearmask >> c) & 1u) {
                if (c & 1)
                  wk[c / 2].v.y = 0.f;
                else
                  wk[c / 2].v.x = 0.f;
                if constexpr (MODE == TC_GRAD) {
                  hi[c] = 0u;
                  lo[c] = 0u;
                }
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) sc2 = padd(sc2, wk[c]);
          if constexpr (MODE == TC_GRAD) {
            tmem_st16(tcol + bt * 16, hi);
            tmem_st16(tcol + NC + bt * 16, lo);
          }
        }
        if constexpr (MODE == TC_GRAD) tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_cc[st]);
      }

      // ---- epilogue: G from TMEM, feature gradient, J_FK^T, records into shared memory ---------------------------
      float gx[DC_MAX_DOF];
#pragma unroll
      for (int i = 0; i < DC_MAX_DOF; ++i) gx[i] = 0.f;
      mbar_wait_wd(bar_g, (uint32_t)((t - t0) & 1));
      tc_fence_after();
      if constexpr (MODE == TC_GRAD) {
        uint32_t gv[16];
        tmem_ld16(tm_lane + L::COL_G, gv);
        tmem_wait_ld();
        if (a.dbg != nullptr && t == 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) a.dbg[(size_t)L::TM * (nch * NC) + tid * 16 + c] = __uint_as_float(gv[c]);
        }
        const float csum = __uint_as_float(gv[L::ONES_ROW]);
#pragma unroll
        for (int f = 0; f < FM; ++f)
          gx[f] = a.rc.grad_scale * (fmaf(x[f], csum, -__uint_as_float(gv[f])) + gex[f]);
      }
      tc_fence_before();
      if (tid < nq) {
        float* rec = os + tid * n_out;
        rec[0] = a.rc.score_scale * ((sc2.lo() + sc2.hi()) + sc_ex);
        if constexpr (MODE == TC_GRAD) {
          const float scale = (a.grad_out != nullptr) ? a.grad_out[b_base + tid] : 1.f;
          if (a.fk.type == DC_FK_NONE) {
            for (int f = 0; f < F; ++f) rec[1 + f] = scale * gx[f];
          } else {
            float xl[DC_MAX_DOF], gq[DC_MAX_DOF];
#pragma unroll
            for (int i = 0; i < DC_MAX_DOF; ++i) {
              xl[i] = 0.f;
              gq[i] = 0.f;
            }
#pragma unroll
            for (int f = 0; f < FM; ++f) xl[f] = x[f];
            fk_vjp<float>(a.fk, qv, xl, 1, gx, 1, gq);
            for (int i = 0; i < a.n_in; ++i) rec[1 + i] = scale * gq[i];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (fused) {
        float* dst = a.score + (size_t)b_base * n_out;
        const int n_words = nq * n_out;
        if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          for (int i = tid; i < n_words / 4; i += L::TM) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(os)[i];
        } else {
          for (int i = tid; i < n_words; i += L::TM) dst[i] = os[i];
        }
      } else {
        if (tid < nq) a.score[(size_t)(b_base + tid) * a.score_ld] = os[tid * n_out];
        if constexpr (MODE == TC_GRAD) {
          for (int i = tid; i < nq * a.n_in; i += L::TM) {
            const int tq = i / a.n_in, c = i - tq * a.n_in;
            a.grad[(size_t)(b_base + tq) * a.grad_ld + c] = os[tq * n_out + 1 + c];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // qs / os are rewritten by the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, L::TMEM_COLS);
  }
}

inline int tc_n_chunks(long long n_sv) { return (int)((n_sv + TcLayout::NC - 1) / TcLayout::NC); }
inline size_t tc_blob_bytes(long long n_sv) { return (size_t)tc_n_chunks(n_sv) * TcLayout::BLOB_BYTES + 16; }

template <int MODE>
int launch_score_tc(TcArgs& a, int num_sms, cudaStream_t stream) {
  using L = TcLayout;
  a.n_tiles = (int)ceil_div64(a.batch, L::TM);
  a.n_chunks = tc_n_chunks(a.n_sv);
  auto kern = score_tc_kernel<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    DC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SM_BYTES));
    attr_set = true;
  }
  const int grid = (int)min((long long)2 * num_sms, (long long)a.n_tiles);
  kern<<<grid, 160, L::SM_BYTES, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc
