// Fused collision-score kernel, tensor-core form (tcgen05 + TMEM + bulk TMA; RQKernel(p = 2), one class, fp32 I/O, F <= 30:
// operand groups of FG = 16 K slots for F <= 14 — the shapes written out below, two CTAs per SM — or FG = 32 for F <= 30,
// one CTA per SM; everything is a template on FG).
//
//   score[b] = sum_n w_n u_bn^2      g_x[b] = -2 gamma sum_n w_n u_bn^3 (x_b - s_n)      u = 1 / (1 + gamma/2 |x_b - s_n|^2)
//
// Same contract as score_tq_kernel (dc_score_tq.cuh) for diffco/kernel_perceptrons.py:362-370 (DiffCo.score with
// diffco/kernel.py:17-29 and a diffco/model.py feature map) and its autograd backward, but the two contractions run on
// the 5th-generation tensor cores instead of the FP32 pipe:
//
//   GEMM1  T [128 x 96]  = A[128 x 48] . B1[96 x 48]^T      T = tau (1 + gamma/2 rho): kernel width and the "+1" are folded
//                                                           into the operands (kind::f16, K = 3 x 16 slots, fp32 in TMEM)
//   GEMM2  G [128 x 32] += CC[128 x 16] . B2[32 x 16]^T     per 8 support vectors, CC = 2^15 u^3 read from TMEM, the
//                                                           weights folded into B2 = [w s | w]
//
// Every fp32 quantity enters as a sum of 11-bit terms in separate K slots (x = xh + xl, s = sh + sl, cc = ch + cl), so
// products carry ~22 bits and are accumulated in fp32.  Power-of-two scales keep every term inside fp16's range and are
// undone exactly in the epilogue.  The expansion |x|^2 + |s|^2 - 2 x.s leaves an ABSOLUTE error in rho that scales with
// the feature norms (the tensor core aligns its addends to the largest one); it only matters where k'(rho) is large, i.e.
// for the few pairs with small rho: pairs under a per-(query, chunk) threshold derived from the error bound are taken
// out of the tensor-core sums and evaluated exactly with direct differences on the FP32 pipe.
//
// CTA = 8 query warps + 4 service warps, two CTAs per SM, persistent over a contiguous range of 128-query tiles.
//   warp 8   (one elected lane) issues every tcgen05.mma: GEMM2(j) then GEMM1(j+2) — the same thread, so the hardware's
//            in-order execution covers the TMEM hazards between them; tcgen05.commit -> mbarriers;
//   warp 9   streams the GEMM1 operand images (9 KB per 96 supports) L2 -> smem, 2-slot ring, 1-D bulk TMA;
//   warp 10  streams the GEMM2 operand images (12 KB), 2-slot ring; the fp32 weights come from L1;
//   query warps: TMEM lane quarter = warp & 3 <-> 32 queries of the tile, column half = warp >> 2.  Per chunk:
//            tcgen05.ld T (48 columns) -> near test -> packed FP32 (FMUL2/FFMA2), one MUFU.RCP per column PAIR -> score in
//            registers, coefficients split, cvt.rn.f16x2, tcgen05.st over the T columns -> arrive.
//   Tiles are software-pipelined across the two halves of the query warps: while the upper half (owners) runs tile t's
//   epilogue (G from TMEM, J_FK^T, records), the lower half already runs forward kinematics and builds the A operand of
//   tile t+1 and enters its chunk loop; G is double-buffered in TMEM for that.
#pragma once

#include <cuda_fp16.h>

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

// Experiment switches (tools/probe builds set them; the defaults are what the library ships).  Measured on B200 at
// BASELINE configs[1] (profiles/r02g_variants.txt): both alternatives lose — the kernel is bound by issue slots and
// dependency latency at 4 warps per scheduler, not by the XU pipe nor by the two halves stalling in lockstep.
#ifndef DC_TC_INTPACK
#define DC_TC_INTPACK 0   // 1: f16 operand words built with integer ops instead of F2FP (frees the XU pipe): 105.7 us vs 92.9 us
#endif
#ifndef DC_TC_PREFETCH
#define DC_TC_PREFETCH 0  // 1: the first two tcgen05.ld of chunk g + 1 are issued at the end of chunk g (before its st-wait /
#endif                    //    arrive / loop top): 99.8 us vs 95.2 us — the two buffers stay live across the loop edge and
                          //    ptxas spills (112 bytes) what it saves in latency (profiles/r02q_variants.txt)
#ifndef DC_TC_FK_UNROLL
#define DC_TC_FK_UNROLL 1  // joints per iteration of the planar FK loop of the tile prologue (code size vs. latency overlap)
#endif
#ifndef DC_TC_HANDOVER
#define DC_TC_HANDOVER 0  // 1: the lower half leaves its end-of-tile queue drain to the owner warps and goes straight to the
#endif                    //    next tile's FK: the CTA's tile-to-tile gap shrinks from 20 k to 13 k cycles, the kernel does not
                          //    get faster (90.0 vs 89.5 us, 2.26 vs 2.23 ms at 2 M queries) — with two CTAs per SM one CTA's gap
                          //    is the other's chance to run its chunk loop alone; SM-level throughput is what binds
                          //    (profiles/r03e_handover.txt)
#ifndef DC_TC_DEPHASE
#define DC_TC_DEPHASE 0   // 1: the owner half starts each tile half a chunk behind the lower half: 93.6 us vs 92.9 us
#endif

namespace dc {

// FG = K slots per operand group = 16 (F <= 14: the BASELINE shape, two CTAs per SM) or 32 (F <= 30: Panda, dual arms,
// URDF robots; one CTA per SM).  Each of the three GEMM1 groups holds FG - 2 feature slots + 2 constant slots.
template <int FG>
struct TcLayoutT {
  static_assert(FG == 16 || FG == 32, "operand group of 16 or 32 K slots");
  static constexpr int TM = 128;   // queries per tile == UMMA M
  static constexpr int NC = 96;    // support vectors per chunk == UMMA N of GEMM1
  static constexpr int GRP = FG;
  static constexpr int FMAX = FG - 2;  // features
  static constexpr int K1 = 3 * FG;    // GEMM1 K slots
  static constexpr int N2 = 2 * FG;    // GEMM2 N: [w s (FMAX), w, 0 | corrections (FMAX), w correction, 0]
  static constexpr int ONES_ROW = FG - 2;
  static constexpr int KS2 = NC / 8;  // GEMM2 instructions per chunk (8 supports x {ch, cl} each)
  static constexpr int CTAS_PER_SM = FG == 16 ? 2 : 1;
  // operand images (bytes).  Global layout: [n_chunks x B1][n_chunks x B2W][trailer]; a B2W record is the GEMM2 image
  // followed by the chunk's fp32 weights [NC] and its max|s|^2.  (Reading the weights through the L1 instead was
  // measured: with 2 x 112 KB of shared memory only 28 KB of L1 remain, 15 % of those loads missed and showed up as
  // the kernel's top stall, profiles/r02o_*.)
  static constexpr int B1_BYTES = NC * K1 * 2;        // [K1/8][NC][8] f16           9216 (FG 16)
  static constexpr int B2_STEP_BYTES = N2 * 16 * 2;   // per 8 supports: [2][N2][8]  1024
  static constexpr int B2_BYTES = KS2 * B2_STEP_BYTES;  // 12288
  static constexpr int OFF_W = B2_BYTES;              // fp32 weights [NC], then the chunk's max |s|^2
  static constexpr int META_S2MAX = NC;               // float index in the weight section
  static constexpr int B2W_BYTES = B2_BYTES + 512;    // 12800
  // trailer floats: 0 max|s|^2 (bits, atomicMax)  1 max |w| max(1, |s_f|) (bits)  2 max |s_f| (bits)
  //                 3 Sa  4 tau c0  5 tau  6 1/(2^15 Sg)  7 gamma  8 valid (1.0 / 0.0)  9 1/(tau c0)
  static constexpr int TRAILER_FLOATS = 16;
  static constexpr int RS1 = 2, RS2 = 2;    // ring slots (both indexed by chunk parity, like the TMEM stages)
  // TMEM columns: two T / CC stages, then the G accumulator(s): two (tile parity) of 32 columns for FG 16, one of 64
  // for FG 32 — GEMM2 of the next tile cannot start before every owner has read G (it waits for all 256 query threads'
  // cc of chunk 0, and the owners only get there after their epilogue), so one accumulator is enough
  static constexpr int COL_STAGE = NC;
  static constexpr int COL_G = 2 * COL_STAGE;  // 192 ..
  static constexpr int G_BUFS = FG == 16 ? 2 : 1;
  static constexpr int TMEM_COLS = 256;
  // near-pair queue entries per warp: large enough that queues are normally drained once, at the end of the tile — a
  // drain in the middle of the chunk loop stalls its warp for ~1500 cycles and, through the chunk barriers, the whole CTA
  static constexpr int QCAP = 160;
  static constexpr int QWARPS = 8;   // query warps: 4 TMEM lane quarters x 2 column halves
  static constexpr int QTHREADS = QWARPS * 32;
  static constexpr int CTRL_WARP = QWARPS;
  static constexpr int THREADS = QTHREADS + 128;
  static constexpr int QS_DOF = FG == 16 ? 8 : 16;   // staged configurations (FK NONE needs none)
  // shared memory (bytes)
  static constexpr int SM_BAR = 0;
  static constexpr int SM_TRAILER = 136;    // the blob's trailer floats 0..11, copied once per CTA (bytes 136 .. 184)
  static constexpr int SM_TILE0 = 184;      // first tile of this CTA (long long): the out-of-line stages read it back
                                            // instead of redoing the 64-bit division (cold code on the tile-to-tile path)
  static constexpr int SM_TMEM_SLOT = 192;
  static constexpr int SM_QHAND = 208;      // int[4]: queue entries each lower-half warp leaves to its owner warp
  static constexpr int SM_RING1 = 256;
  static constexpr int SM_RING2 = SM_RING1 + RS1 * B1_BYTES;
  static constexpr int SM_A = SM_RING2 + RS2 * B2W_BYTES;   // A [K1/8][128][8] f16
  static constexpr int SM_XS = SM_A + TM * K1 * 2;          // features [2][128][FG] f32 (near-pair path, epilogue)
  static constexpr int XLO_SCALE_LOG2 = 22;                 // low parts are kept as f16 of 2^22 lo (|lo| <= ulp(hi) / 2)
  // exact-path accumulators: score [8][32], feature gradient [8][32][FG] (non-owner warps first); the owners' half
  // (+ 512 bytes) doubles as the output records [128][dof + 1] once the owners have read it
  static constexpr int SM_XLO = SM_XS + 2 * TM * FG * 4;    // low parts of the features [2][128][FG] f16 (near-pair path)
  static constexpr int SM_ACC = SM_XLO + 2 * TM * FG * 2;
  static constexpr int ACC_BYTES = QWARPS * 32 * 4 + QWARPS * 32 * FG * 4 + 512;
  static constexpr int SM_QS = SM_ACC + ACC_BYTES;          // staged configurations [2][128][QS_DOF]
  static constexpr int SM_ROWS = SM_QS + 2 * TM * QS_DOF * 4;
  static constexpr int SM_QUEUE = SM_ROWS + 5 * TM * 4;   // lower-half partial scores [128], near-threshold line [2][128] x 2
  static constexpr int SM_BYTES = SM_QUEUE + QWARPS * QCAP * 4;
  static_assert(CTAS_PER_SM * (SM_BYTES + 1024) <= 233472, "the CTAs of one SM must fit in 228 KB of shared memory");
  static_assert(SM_BYTES <= 232448, "at most 227 KB of dynamic shared memory per CTA");
  static_assert(TM * (DC_MAX_DOF + 1) * 4 <= 4 * 32 * FG * 4 + 512, "output records must fit the owners' accumulators");
  static_assert(TM * (FMAX + 1) * 4 <= 4 * 32 * FG * 4 + 512, "... also for transform=None, where a record is [score | d/dx (F)]");
};
using TcLayout = TcLayoutT<16>;
__host__ __device__ constexpr int tc_group(int n_features) { return n_features <= 14 ? 16 : 32; }

struct TcArgs {
  dc_fk_desc fk;
  RadialConsts<float> rc;
  const unsigned char* blob;  // [n_chunks x B1][n_chunks x B2W][trailer]
  const float* table;         // packed [-s | w] rows (dc_pack_supports), for the exact near-pair path
  const float* table_lo;      // optional: low parts -(s - fl32(s)) in the same row layout (dc_pack_supports_lo)
  const float* q;
  float* score;
  float* grad;
  const float* grad_out;
  long long* trace;  // optional clock64 timeline (tools/probe/tc_probe.cu), 16 slots per chunk
  float* dbg;        // optional debug dump (tools/probe/tc_probe.cu): tile 0 rho [128][n_chunks*NC] then G [128][32], x [128][16]
  unsigned long long* stats;  // optional: [0] += near pairs evaluated exactly
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
  float* mirror;               // optional: this launch's records are ALSO written here, row 0 = first query of the launch
                               // (a mapped pinned host buffer: the end-to-end path of the multi-GPU scorer)
  float err_coef;  // delta(rho) <= err_coef * (|x|^2 + max_chunk |s|^2)
  float tol_pair;  // admissible |w|-relative error of one pair
  // Step barrier folded into the kernel's tail (dc_score_grad_bcast_sync; sync_world == 0: none).  The last CTA to finish
  // (counted in *done) publishes `epoch` into this rank's slot of every rank's flag array and waits for every rank's epoch
  // in its own — what dc_peer_barrier does as a second launch.
  uint32_t* flag_peer[DC_MAX_PEERS];  // my slot in rank r's flag array
  const uint32_t* flag_mine;          // my flag array (slot r = rank r's epoch)
  unsigned int* done;                 // device counter, 0 between launches
  long long sync_timeout_cycles;
  uint32_t epoch;
  int sync_world;
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
// The waiting thread is suspended by the hardware until the phase completes or the time hint (ns) expires, so a waiting
// warp issues a handful of instructions per microsecond instead of spinning: with the default (short, system-dependent)
// time limit the four service warps' spin loops were 23 % of all instructions the kernel issued and competed with the
// query warps for issue slots (profiles/r02y_score_tc_bench_hot_sass.txt: 1.75 M loop iterations x 8 instructions per launch).
#ifndef DC_TC_WAIT_HINT_NS
#define DC_TC_WAIT_HINT_NS 20000
#endif
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_test(addr, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((uint32_t)DC_TC_WAIT_HINT_NS)
        : "memory");
    if (done) return;
    if ((++spins & 1023u) == 0) {  // watchdog, off the fast path: ~30 s means a protocol bug, not a slow launch
      if (t0 == 0)
        t0 = clock64();
      else if (clock64() - t0 > 60000000000LL)
        __trap();
    }
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

__host__ __device__ inline int tc_n_chunks(long long n_sv) { return (int)((n_sv + TcLayout::NC - 1) / TcLayout::NC); }
// fg: the operand group size the image was packed for (tc_group(n_features))
__host__ __device__ inline size_t tc_trailer_offset(long long n_sv, int fg = 16) {
  const size_t per_chunk = fg == 16 ? (size_t)(TcLayoutT<16>::B1_BYTES + TcLayoutT<16>::B2W_BYTES)
                                    : (size_t)(TcLayoutT<32>::B1_BYTES + TcLayoutT<32>::B2W_BYTES);
  return (size_t)tc_n_chunks(n_sv) * per_chunk;
}
__host__ __device__ inline size_t tc_blob_bytes(long long n_sv, int fg = 16) {
  return tc_trailer_offset(n_sv, fg) + TcLayout::TRAILER_FLOATS * 4;
}
__host__ __device__ inline const float* tc_trailer(const void* blob, long long n_sv, int fg = 16) {
  return reinterpret_cast<const float*>(static_cast<const unsigned char*>(blob) + tc_trailer_offset(n_sv, fg));
}

// ---- pack: support vectors -> operand images ----------------------------------------------------------------------
// trailer[0..2] = max |s|^2, max |w| max(1, max_f |s_f|), max |s_f| (as int bit patterns; zeroed by the caller)
__global__ void __launch_bounds__(128) tc_scan_kernel(const float* __restrict__ s, const float* __restrict__ w, int n, int F,
                                                        int* __restrict__ trailer) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ss = 0.f, sm = 0.f;
  for (int f = 0; f < F; ++f) {
    const float v = s[(size_t)i * F + f];
    ss = fmaf(v, v, ss);
    sm = fmaxf(sm, fabsf(v));
  }
  atomicMax(&trailer[0], __float_as_int(ss));
  atomicMax(&trailer[1], __float_as_int(fabsf(w[i]) * fmaxf(1.f, sm)));
  atomicMax(&trailer[2], __float_as_int(sm));
}

// Scales of the operand images for RQKernel(gamma, p = 2).
//   T = tau (1 + c0 rho), c0 = gamma / 2, tau = 2^-5: u' = 1/T <= 32, k' = u'^2 = 2^10 u^2, cc = u'^3 = 2^15 u^3 <= 2^15 (f16).
//   A side holds Sa x, B side (-2 tau c0 / Sa) s with Sa the power of two nearest sqrt(2 tau c0): both sides the same size.
//   GEMM2's B operand holds Sg w [s | 1] with Sg a power of two, Sg max|w s| <= 2^8.
struct TcScales {
  float sa, tc0, tau, beta, sg, inv_g;
  bool valid;
};
__device__ __forceinline__ TcScales tc_scales(const float* trailer, float gamma) {
  TcScales c;
  const float ssmax = trailer[0], wsmax = trailer[1], sfmax = trailer[2];
  c.tau = 1.f / 32.f;
  c.tc0 = c.tau * 0.5f * gamma;
  c.sa = pow2_floor(sqrtf(2.f * c.tc0) * 1.41421356f);
  c.sa = fminf(fmaxf(c.sa, 1.f / 4096.f), 4096.f);
  c.beta = -2.f * c.tc0 / c.sa;
  c.sg = (wsmax > 0.f) ? pow2_floor(256.f / wsmax) : 1.f;
  c.sg = fminf(fmaxf(c.sg, 1.f / 1048576.f), 1048576.f);
  c.inv_g = 1.f / (32768.f * c.sg);
  // every 11-bit term of the dominant features (and its /256, x256 companions) must be a normal f16 number
  const float am = c.sa * sfmax, bm = fabsf(c.beta) * sfmax;
  c.valid = gamma > 0.f && am <= 8192.f && bm <= 8192.f && am >= 1.f / 64.f && bm >= 1.f / 64.f &&
            (c.tau + c.tc0 * ssmax) <= 30000.f && wsmax > 0.f && c.sg * wsmax <= 512.f;
  return c;
}

// One thread per (chunk, local support index).  s_feat[N, F] are the transformed supports, w[N] the weights.  The whole
// blob was zeroed by the caller (the chunk maxima are accumulated with atomicMax).
template <int FG>
__global__ void __launch_bounds__(128) pack_supports_tc_kernel(const float* __restrict__ s, const float* __restrict__ w,
                                                                 int n, int F, int n_chunks, float gamma,
                                                                 unsigned char* __restrict__ blob) {
  using L = TcLayoutT<FG>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks * L::NC) return;
  unsigned char* blob2 = blob + (size_t)n_chunks * L::B1_BYTES;
  float* trailer = reinterpret_cast<float*>(blob2 + (size_t)n_chunks * L::B2W_BYTES);
  const TcScales sc = tc_scales(trailer, gamma);
  if (idx == 0) {
    trailer[3] = sc.sa;
    trailer[4] = sc.tc0;
    trailer[5] = sc.tau;
    trailer[6] = sc.inv_g;
    trailer[7] = gamma;
    trailer[8] = sc.valid ? 1.f : 0.f;
    trailer[9] = 1.f / sc.tc0;
  }
  const int j = idx / L::NC, r = idx - j * L::NC;
  float b1[L::K1], mainv[FG], corrv[FG];
#pragma unroll
  for (int k = 0; k < L::K1; ++k) b1[k] = 0.f;
#pragma unroll
  for (int k = 0; k < FG; ++k) mainv[k] = corrv[k] = 0.f;
  float wv = 0.f;
  float* w_sec = reinterpret_cast<float*>(blob2 + (size_t)j * L::B2W_BYTES + L::OFF_W);
  if (idx < n) {
    wv = w[idx];
    const float gw = sc.sg * wv;  // exact (power of two)
    float ss = 0.f;
#pragma unroll
    for (int f = 0; f < L::FMAX; ++f) {
      if (f < F) {
        const float sv = s[(size_t)idx * F + f];
        const float m = sc.beta * sv;  // (-2 tau c0 / Sa) s
        const float hi = split_hi(m);
        b1[f] = hi;                          // pairs with xh
        b1[FG + f] = hi * (1.f / 256.f);     // pairs with 256 xl
        b1[2 * FG + f] = (m - hi) * 256.f;   // pairs with xh / 256
        const float p = gw * sv;
        const float ph = split_hi(p);
        mainv[f] = ph;
        corrv[f] = p - ph;
        ss = fmaf(sv, sv, ss);
      }
    }
    const float S = fmaf(sc.tc0, ss, sc.tau);  // tau (1 + c0 |s|^2)
    const float s1 = split_hi(S), s2 = split_hi(S - s1), s3 = S - s1 - s2;
    b1[FG - 2] = s1;       // x 1
    b1[FG - 1] = s2;       // x 1
    b1[2 * FG - 2] = s3;   // x 1
    b1[2 * FG - 1] = 1.f;  // x X1   (X = tau c0 |x|^2, three terms)
    b1[3 * FG - 2] = 1.f;  // x X2
    b1[3 * FG - 1] = 1.f;  // x X3
    const float gh = split_hi(gw);
    mainv[L::ONES_ROW] = gh;
    corrv[L::ONES_ROW] = gw - gh;
    atomicMax(reinterpret_cast<int*>(w_sec) + L::META_S2MAX, __float_as_int(ss));
  } else {
    b1[FG - 2] = 32768.f;  // padding rows: T >= 2^15, never near; weight 0 removes them from every sum
  }
  __half* b1p = reinterpret_cast<__half*>(blob + (size_t)j * L::B1_BYTES);
#pragma unroll
  for (int k = 0; k < L::K1; ++k) b1p[(k >> 3) * (L::NC * 8) + r * 8 + (k & 7)] = __float2half_rn(b1[k]);
  // GEMM2 image of the 8-support step ks = r / 8: K slot i = r % 8 multiplies ch of support r, slot 8 + i its cl
  __half* b2p = reinterpret_cast<__half*>(blob2 + (size_t)j * L::B2W_BYTES + (r >> 3) * L::B2_STEP_BYTES);
  const int i = r & 7;
#pragma unroll
  for (int f = 0; f < FG; ++f) {
    const __half hm = __float2half_rn(mainv[f]);
    b2p[0 * (L::N2 * 8) + f * 8 + i] = hm;                                // slot i,     row f:      (Sg w s)_h (row FG - 2: (Sg w)_h)
    b2p[1 * (L::N2 * 8) + f * 8 + i] = hm;                                // slot 8 + i, row f
    b2p[0 * (L::N2 * 8) + (FG + f) * 8 + i] = __float2half_rn(corrv[f]);  // slot i,     row FG + f: low parts
    b2p[1 * (L::N2 * 8) + (FG + f) * 8 + i] = __float2half_rn(0.f);
  }
  w_sec[r] = wv;
}

#ifdef DC_TC_ENABLE_TRACE
#define DC_TC_TRACE(slot, gidx) \
  do { if (a.trace != nullptr && blockIdx.x == 1 && lane == 0 && (gidx) < 64) a.trace[(gidx) * 16 + (slot)] = clock64(); } while (0)
#else
#define DC_TC_TRACE(slot, gidx) do { } while (0)
#endif
// tile events: slot = local tile index (< 16), row 50 + ev (non-owner warp 0: ev 0..4, owner warp 4: ev 8..12)
#define DC_TC_TRACE_TILE(w, ev) do { if (warp == (w)) DC_TC_TRACE(ti, 50 + (ev)); } while (0)

// ---- the kernel -----------------------------------------------------------------------------------------------
enum TcMode { TC_SCORE = 0, TC_GRAD = 1 };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// named barriers (0 is __syncthreads)
enum { TCB_FK = 1, TCB_EPI = 2, TCB_OWN = 3, TCB_LOW = 4, TCB_ACC = 5, TCB_END = 6, TCB_PHASE = 7 };

// Exact evaluation of queued near pairs, one pair per lane: the lane reads its query's features (shared memory) and the
// fp32 support row [-s | w] (global; prefetched into L1 when the pair was queued; all loads of the batch are in flight
// together), evaluates the pair with direct differences — the same arithmetic as the FP32-pipe kernels — and adds its
// score / feature-gradient terms to the WARP'S OWN shared-memory accumulators of that row.  Entries of one row are
// applied strictly in queue (= column) order: lanes holding the same row take turns (match.any groups them; almost
// always a single round), so a row's result depends neither on its position in the batch nor on timing.
// (Out of line, and it re-derives its shared-memory pointers from (warp, tile parity): the chunk loop that calls it keeps
// as little state alive across the call as possible.)
template <int FG>
__device__ __noinline__ void tc_drain_pairs(const TcArgs& a, int count, int warp, int buf) {
  using L = TcLayoutT<FG>;
  constexpr int FM = L::FMAX;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const uint32_t* queue = reinterpret_cast<const uint32_t*>(smem + L::SM_QUEUE) + warp * L::QCAP;
  const float* xs = reinterpret_cast<const float*>(smem + L::SM_XS) + buf * L::TM * FG;
  const __half* xlo = reinterpret_cast<const __half*>(smem + L::SM_XLO) + buf * L::TM * FG;
  float* sacc_w = reinterpret_cast<float*>(smem + L::SM_ACC) + warp * 32;
  float* gacc_w = reinterpret_cast<float*>(smem + L::SM_ACC) + L::QWARPS * 32 + warp * 32 * FG;
  const float lo_scale = 1.f / (float)(1 << L::XLO_SCALE_LOG2);
  for (int head = 0; head < count; head += 32) {
    const bool valid = head + lane < count;
    const uint32_t e = queue[valid ? head + lane : head];
    const int row = (int)(e >> 24), n = (int)(e & 0xffffffu);
    const float4* tr = reinterpret_cast<const float4*>(a.table + (size_t)n * a.row_stride);  // 16-byte aligned rows
    const float4* xr = reinterpret_cast<const float4*>(xs + row * FG);
    float t[FG], x[FG];
#pragma unroll
    for (int v = 0; v < FG / 4; ++v) {
      const float4 tv = (4 * v < a.row_stride) ? tr[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xv = xr[v];
      t[4 * v] = tv.x, t[4 * v + 1] = tv.y, t[4 * v + 2] = tv.z, t[4 * v + 3] = tv.w;
      x[4 * v] = xv.x, x[4 * v + 1] = xv.y, x[4 * v + 2] = xv.z, x[4 * v + 3] = xv.w;
    }
    float wv = 0.f;
#pragma unroll
    for (int f = 0; f < FG; ++f) wv = (f == a.f_pad) ? t[f] : wv;  // the weight sits right after the padded features
    // low parts: x = x_hi + x_lo and s = s_hi + s_lo from the float64 feature map, so that the difference is good to ~1e-9
    // even when the pair is 1e-3 apart (a float32 feature alone is off by up to half an ulp of ITS magnitude)
    float lo[FG];
    {
      const uint4* lr = reinterpret_cast<const uint4*>(xlo + row * FG);
#pragma unroll
      for (int q4 = 0; q4 < FG / 8; ++q4) {
        const uint4 l = lr[q4];
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&lw[v]));
          lo[8 * q4 + 2 * v] = p.x * lo_scale;
          lo[8 * q4 + 2 * v + 1] = p.y * lo_scale;
        }
      }
    }
    if (a.table_lo != nullptr) {
      const float4* tl = reinterpret_cast<const float4*>(a.table_lo + (size_t)n * a.row_stride);
#pragma unroll
      for (int v = 0; v < FG / 4; ++v) {
        const float4 tv = (4 * v < a.row_stride) ? tl[v] : make_float4(0.f, 0.f, 0.f, 0.f);
        lo[4 * v] += tv.x, lo[4 * v + 1] += tv.y, lo[4 * v + 2] += tv.z, lo[4 * v + 3] += tv.w;
      }
    }
    float d[FG], rho = 0.f;
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      d[f] = (f < FM && f < a.n_feat) ? (x[f] + t[f]) + lo[f] : 0.f;
      rho = fmaf(d[f], d[f], rho);
    }
    float k, coef;
    radial_eval<KR_RQ2, float>(a.rc, rho, k, coef);
    const float cg = wv * coef, cs = wv * k;
    const int r = row & 31;
    // lanes with the same row: the one with the fewest lower-numbered peers goes first
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? r : 32 + lane);
    int turn = __popc(peers & ((1u << lane) - 1u));
    uint32_t pending = __ballot_sync(0xffffffffu, valid);
    while (pending != 0) {
      if (valid && turn == 0) {
        float4* g4 = reinterpret_cast<float4*>(gacc_w + r * FG);
#pragma unroll
        for (int v = 0; v < FG / 4; ++v) {
          float4 g = g4[v];
          g.x = fmaf(cg, d[4 * v + 0], g.x);  // d is zero past the features
          g.y = fmaf(cg, d[4 * v + 1], g.y);
          g.z = fmaf(cg, d[4 * v + 2], g.z);
          g.w = fmaf(cg, d[4 * v + 3], g.w);
          g4[v] = g;
        }
        sacc_w[r] += cs;
      }
      --turn;
      __syncwarp();
      pending = __ballot_sync(0xffffffffu, valid && turn >= 0);
    }
  }
  if (a.stats != nullptr && lane == 0) atomicAdd(a.stats, (unsigned long long)count);
}

// ---- lower half of the query warps: configurations of tile ti -> FK (float64) -> features (hi, lo), |x|^2, A operand ----------
// Out of line on purpose: its register needs (float64 sincos chain) must not leak into the allocation of the chunk loop.
// It re-derives what it needs (tile range, scales) instead of taking it from the caller, for the same reason.
template <int FG>
__device__ __noinline__ void tc_fk_stage(const TcArgs& a, int ti, int tid) {
  using L = TcLayoutT<FG>;
  constexpr int NC = L::NC, FM = L::FMAX, QT = L::QTHREADS, TM = L::TM;
  (void)NC;
  extern __shared__ __align__(128) unsigned char smem[];
  const int row = tid, lane = tid & 31, warp = tid >> 5;
  (void)lane;
  (void)warp;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(smem + L::SM_BAR) + 10;
  uint64_t* bar_qfull = reinterpret_cast<uint64_t*>(smem + L::SM_BAR) + 13;
  unsigned char* a_op = smem + L::SM_A;
  float* xs_all = reinterpret_cast<float*>(smem + L::SM_XS);
  __half* xlo_all = reinterpret_cast<__half*>(smem + L::SM_XLO);
  float* qs_all = reinterpret_cast<float*>(smem + L::SM_QS);
  float2* thr_all = reinterpret_cast<float2*>(smem + L::SM_ROWS + TM * 4);
  const long long t0 = *reinterpret_cast<const long long*>(smem + L::SM_TILE0);
#ifdef DC_TC_ENABLE_TRACE
  const int nch = a.n_chunks;
#endif
  DC_TC_TRACE_TILE(0, 7);
  const float* trailer = reinterpret_cast<const float*>(smem + L::SM_TRAILER);
  const float sa = trailer[3], tc0 = trailer[4];
  const int F = a.n_feat;
  const bool has_fk = a.fk.type != DC_FK_NONE;
  const int buf = ti & 1;
  const long long b_base = (t0 + ti) * TM;
  const int nq = (int)min((long long)TM, a.batch - b_base);
  float* xs = xs_all + buf * TM * FG;
  float* qs = qs_all + buf * TM * L::QS_DOF;
  const float* src = a.q + (size_t)b_base * a.n_in;
  const int n_words = nq * a.n_in;
  float xx = 0.f, xamax = 0.f;
  bool in_range;
  const float xlo_sc = (float)(1 << L::XLO_SCALE_LOG2);
  if (FG == 16 && has_fk && a.fk.type == DC_FK_PLANAR_CHAIN && a.fk.n_repeat <= 1 && !a.fk.time_last) {
    // The BASELINE robot.  One ROLLED loop over the joints that writes features, low parts and the A-operand words
    // straight to shared memory: this stage runs once per tile on four warps, its code is cold every time, and as
    // 1150 unrolled instructions it spent half its cycles waiting for instruction fetches (ncu: stall_no_inst, 53 % of
    // the stage's samples).  The arithmetic per joint is fk_planar_f64_reg's, operation for operation, so the features
    // are bit-identical to dc_fk_forward's (support_transformed) and to every other kernel's.
    mbar_wait_wd(bar_qfull + buf, (uint32_t)((ti >> 1) & 1));
    DC_TC_TRACE_TILE(0, 4);
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int kc = 0; kc < L::K1 / 8; ++kc) reinterpret_cast<uint4*>(a_op)[kc * TM + row] = z4;
    {
      uint4* xr = reinterpret_cast<uint4*>(xs + row * FG);
#pragma unroll
      for (int v = 0; v < FG / 4; ++v) xr[v] = z4;
      uint4* lr = reinterpret_cast<uint4*>(xlo_all + (buf * TM + row) * FG);
#pragma unroll
      for (int v = 0; v < FG / 8; ++v) lr[v] = z4;
    }
    const int nl = (row < nq) ? a.fk.n_links : 0;
    const float* qr = qs + row * a.n_in;
    unsigned char* arow = a_op + row * 16;
    double th = 0.0, px = 0.0, py = 0.0;
    constexpr int kFkUnroll = DC_TC_FK_UNROLL;
#pragma unroll kFkUnroll
    for (int i = 0; i < nl; ++i) {
      th += (double)qr[i];
      double sn, cs;
      sincos_fast64(th, &sn, &cs);
      const double len = a.fk.link_length[i];
      px = fma(len, cs, px);
      py = fma(len, sn, py);
      const float hx = (float)px, hy = (float)py;
      const float lx = (float)(px - (double)hx), ly = (float)(py - (double)hy);
      *reinterpret_cast<float2*>(xs + row * FG + 2 * i) = make_float2(hx, hy);
      reinterpret_cast<uint32_t*>(xlo_all + (buf * TM + row) * FG)[i] = pack_f16x2(lx * xlo_sc, ly * xlo_sc);
      xx = fmaf(hx, hx, xx);
      xx = fmaf(hy, hy, xx);
      xamax = fmaxf(xamax, fmaxf(fabsf(hx), fabsf(hy)));
      // K slots 2i, 2i + 1 (x beta s_h), 16 + .. (x beta s_h / 256), 32 + .. (x 256 (beta s)_lo)
      const float ax = sa * hx, ay = sa * hy;
      const float hix = split_hi(ax), hiy = split_hi(ay);
      unsigned char* aw = arow + ((2 * i) >> 3) * (TM * 16) + ((2 * i) & 7) * 2;
      *reinterpret_cast<uint32_t*>(aw) = pack_f16x2(hix, hiy);
      *reinterpret_cast<uint32_t*>(aw + (FG / 8) * (TM * 16)) = pack_f16x2((ax - hix) * 256.f, (ay - hiy) * 256.f);
      *reinterpret_cast<uint32_t*>(aw + (2 * FG / 8) * (TM * 16)) = pack_f16x2(hix * (1.f / 256.f), hiy * (1.f / 256.f));
    }
    DC_TC_TRACE_TILE(0, 5);
    // queries whose scaled features leave f16's range take the exact path for every pair (A row zeroed)
    in_range = (sa * xamax < 32768.f) && (tc0 * xx < 32768.f);
    if (!in_range) {
#pragma unroll
      for (int kc = 0; kc < L::K1 / 8; ++kc) reinterpret_cast<uint4*>(a_op)[kc * TM + row] = z4;
    }
    const float XX = in_range ? tc0 * xx : 0.f;
    const float x1 = split_hi(XX), x2 = split_hi(XX - x1), x3 = XX - x1 - x2;
    // the two constant slots that close each group of FG: x S1, x S2 | x S3, x 1 | x 1, x 1
    *reinterpret_cast<uint32_t*>(arow + (FG / 8 - 1) * (TM * 16) + 12) = pack_f16x2(1.f, 1.f);
    *reinterpret_cast<uint32_t*>(arow + (2 * FG / 8 - 1) * (TM * 16) + 12) = pack_f16x2(1.f, x1);
    *reinterpret_cast<uint32_t*>(arow + (3 * FG / 8 - 1) * (TM * 16) + 12) = pack_f16x2(x2, x3);
  } else {
    float x[FM], xlo[FM];  // features as float32 (hi, lo) pairs of a float64 evaluation (dc_fk.cuh: fk_forward_f32x)
    float qv[DC_MAX_DOF];
    if (has_fk) {
      // the tile's configurations were staged (coalesced) by the prefetch warp while the previous tile was being scored
      mbar_wait_wd(bar_qfull + buf, (uint32_t)((ti >> 1) & 1));
      DC_TC_TRACE_TILE(0, 4);
#pragma unroll
      for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (row < nq && i < a.n_in) ? qs[row * a.n_in + i] : 0.f;
      {
        float xh16[FG], xl16[FG];
#pragma unroll
        for (int i = 0; i < FG; ++i) xh16[i] = xl16[i] = 0.f;
        if (row < nq) fk_forward_f32x<true>(a.fk, qv, xh16, 1, xl16, 1);
#pragma unroll
        for (int f = 0; f < FM; ++f) {
          x[f] = (f < F) ? xh16[f] : 0.f;
          xlo[f] = (f < F) ? xl16[f] : 0.f;
        }
      }
    } else {
      // transform=None: the rows ARE the features; staged straight into the feature layout [128][16]
      for (int i = tid; i < TM * FG; i += TM) xs[i] = 0.f;
      named_sync(TCB_LOW, TM);
      for (int i = tid; i < n_words; i += TM) {
        const int r = i / a.n_in;
        xs[r * FG + (i - r * a.n_in)] = src[i];
      }
      named_sync(TCB_LOW, TM);
#pragma unroll
      for (int f = 0; f < FM; ++f) {
        x[f] = xs[row * FG + f];
        xlo[f] = 0.f;
      }
    }
    DC_TC_TRACE_TILE(0, 5);
    {
      const float sc = (float)(1 << L::XLO_SCALE_LOG2);
      uint4* lr = reinterpret_cast<uint4*>(xlo_all + (buf * TM + row) * FG);
      auto lo_at = [&](int f) { return f < FM ? xlo[f < FM ? f : 0] * sc : 0.f; };
#pragma unroll
      for (int v = 0; v < FG / 8; ++v) {
        uint4 l;
        l.x = pack_f16x2(lo_at(8 * v + 0), lo_at(8 * v + 1)), l.y = pack_f16x2(lo_at(8 * v + 2), lo_at(8 * v + 3));
        l.z = pack_f16x2(lo_at(8 * v + 4), lo_at(8 * v + 5)), l.w = pack_f16x2(lo_at(8 * v + 6), lo_at(8 * v + 7));
        lr[v] = l;
      }
    }
#pragma unroll
    for (int f = 0; f < FM; ++f) {
      xx = fmaf(x[f], x[f], xx);
      xamax = fmaxf(xamax, fabsf(x[f]));
    }
    {
      float4* xr = reinterpret_cast<float4*>(xs + row * FG);
      auto x_at = [&](int f) { return f < FM ? x[f < FM ? f : 0] : 0.f; };
#pragma unroll
      for (int v = 0; v < FG / 4; ++v) xr[v] = make_float4(x_at(4 * v), x_at(4 * v + 1), x_at(4 * v + 2), x_at(4 * v + 3));
    }
    // queries whose scaled features leave f16's range take the exact path for every pair (A row zeroed)
    in_range = (sa * xamax < 32768.f) && (tc0 * xx < 32768.f);
    {
      float v[L::K1];
      const float XX = in_range ? tc0 * xx : 0.f;
#pragma unroll
      for (int k = 0; k < L::K1; ++k) v[k] = 0.f;
#pragma unroll
      for (int f = 0; f < FM; ++f) {
        const float xsc = in_range ? sa * x[f] : 0.f;
        const float hi = split_hi(xsc);
        v[f] = hi;                           // x beta s_h
        v[FG + f] = (xsc - hi) * 256.f;      // x beta s_h / 256
        v[2 * FG + f] = hi * (1.f / 256.f);  // x 256 (beta s)_lo
      }
      const float x1 = split_hi(XX), x2 = split_hi(XX - x1), x3 = XX - x1 - x2;
      v[FG - 2] = 1.f;      // x S1   (S = tau (1 + c0 |s|^2), three terms)
      v[FG - 1] = 1.f;      // x S2
      v[2 * FG - 2] = 1.f;  // x S3
      v[2 * FG - 1] = x1;   // x 1
      v[3 * FG - 2] = x2;   // x 1
      v[3 * FG - 1] = x3;   // x 1
#pragma unroll
      for (int kc = 0; kc < L::K1 / 8; ++kc) {
        uint4 pk;
        pk.x = pack_f16x2(v[8 * kc + 0], v[8 * kc + 1]);
        pk.y = pack_f16x2(v[8 * kc + 2], v[8 * kc + 3]);
        pk.z = pack_f16x2(v[8 * kc + 4], v[8 * kc + 5]);
        pk.w = pack_f16x2(v[8 * kc + 6], v[8 * kc + 7]);
        reinterpret_cast<uint4*>(a_op)[kc * TM + row] = pk;
      }
    }
  }
  fence_proxy_async();
  mbar_arrive(bar_a);
  DC_TC_TRACE_TILE(0, 6);
  {
    // near threshold of this query on T as a line in the chunk's max|s|^2: pairs with 1 + c0 rho < (gamma drho / tol)^(1/3),
    // drho = err (|x|^2 + max_chunk |s|^2), are recomputed exactly.  The cube root is bounded from above by its tangent at the
    // largest chunk maximum (concave function): exact for the widest chunk, conservative for the others, one FFMA per chunk.
    const float s2max = trailer[0], tau = trailer[5];
    const float kq = -a.rc.grad_scale * 0.5f * a.err_coef / a.tol_pair;  // gamma err / tol
    const float base = fmaxf(kq * (xx + s2max), 1.f);
    const float cr = exp2f(__log2f(base) * (1.f / 3.f));  // two MUFU ops; the 0.1 % margin below covers their error
    const float c1 = tau * 1.001f * kq / (3.f * cr * cr);
    const float c0 = in_range ? tau * 1.001f * cr - c1 * s2max : 3.0e38f;  // out-of-range rows: every pair is "near"
    thr_all[buf * TM + row] = make_float2(c0, c1);
  }
#ifdef DC_TC_ENABLE_TRACE
  if (a.dbg != nullptr && t0 + ti == 0) {
#pragma unroll
    for (int c = 0; c < 16; ++c) a.dbg[(size_t)TM * (nch * NC) + TM * 32 + row * 16 + c] = c < FM ? xs[row * FG + c] : 0.f;
  }
#endif
  named_arrive(TCB_FK, QT);  // features, |x|^2 (and the staged configurations) of tile ti are visible to the owners
}

// ---- epilogue of tile ti (owner half of the query warps): G from TMEM, feature gradient, J_FK^T, records -> memory -----------
// Out of line like tc_fk_stage: its register arrays must not shape the allocation of the chunk loop.  `sc_hi` is the
// caller's partial score (upper column half); everything else is re-derived.
template <int MODE, int FG>
__device__ __noinline__ void tc_epilogue(const TcArgs& a, int ti, int ntile, float sc_hi) {
  using L = TcLayoutT<FG>;
  constexpr int NC = L::NC, FM = L::FMAX, QT = L::QTHREADS, TM = L::TM;
  (void)NC;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = ((warp & 3) << 5) | lane;
  const int buf = ti & 1;
  uint64_t* bar_g = reinterpret_cast<uint64_t*>(smem + L::SM_BAR) + 11;
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(smem + L::SM_TMEM_SLOT);
  const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float* xs = reinterpret_cast<const float*>(smem + L::SM_XS) + buf * TM * FG;
  float* sacc = reinterpret_cast<float*>(smem + L::SM_ACC);
  float* gacc = sacc + L::QWARPS * 32;
  float* os = gacc + 4 * 32 * FG;
  float* sacc_w = sacc + warp * 32;
  float* gacc_w = gacc + warp * 32 * FG;
  float* qs_all = reinterpret_cast<float*>(smem + L::SM_QS);
  const float* sc_p = reinterpret_cast<const float*>(smem + L::SM_ROWS);
  const long long t0 = *reinterpret_cast<const long long*>(smem + L::SM_TILE0);
  const long long b_base = (t0 + ti) * TM;
  const int nq = (int)min((long long)TM, a.batch - b_base);
  const int F = a.n_feat;
  const int n_out = 1 + (MODE == TC_GRAD ? a.n_in : 0);
  const bool fused = (MODE == TC_GRAD) ? (a.score_ld == a.grad_ld && a.score_ld == n_out && a.grad == a.score + 1)
                                       : (a.score_ld == 1);
  const bool has_fk = a.fk.type != DC_FK_NONE;
  const float inv_g = reinterpret_cast<const float*>(smem + L::SM_TRAILER)[6];
#ifdef DC_TC_ENABLE_TRACE
  const int nch = a.n_chunks;
#endif
  const P2 sc2(sc_hi, 0.f);
  // ---- epilogue (owners): G from TMEM, feature gradient, J_FK^T, records into shared memory ------------------------
  named_sync(TCB_EPI, QT);
  {
    // entries the lower-half warp of this lane quarter handed over (tiles with a successor): applied to ITS accumulators
    const int handed = reinterpret_cast<const int*>(smem + L::SM_QHAND)[warp - 4];
    if (handed > 0) {
      tc_drain_pairs<FG>(a, handed, warp - 4, buf);
      __syncwarp();
    }
  }
  float gx[FG], xl[FG];
#pragma unroll
  for (int i = 0; i < FG; ++i) {
    gx[i] = 0.f;
    xl[i] = xs[row * FG + i];
  }
  // G accumulator(s): two of them alternate with the tile parity (FG 16), or one is reused by every tile (FG 32)
  const int gb = (L::G_BUFS == 2) ? (ti & 1) : 0;
  mbar_wait_wd(&bar_g[gb], (uint32_t)(((L::G_BUFS == 2) ? (ti >> 1) : ti) & 1));
  tc_fence_after();
  DC_TC_TRACE_TILE(4, 11);
  if constexpr (MODE == TC_GRAD) {
    uint32_t gm[FG], gc[FG];  // columns [0, FG): sum cc w [s | 1] (hi terms); [FG, 2 FG): the low-part corrections
#pragma unroll
    for (int v = 0; v < FG / 16; ++v) {
      tmem_ld16(tm_lane + L::COL_G + (uint32_t)gb * L::N2 + 16 * v, gm + 16 * v);
      tmem_ld16(tm_lane + L::COL_G + (uint32_t)gb * L::N2 + FG + 16 * v, gc + 16 * v);
    }
    tmem_wait_ld();
#ifdef DC_TC_ENABLE_TRACE
    if (a.dbg != nullptr && t0 + ti == 0) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {  // (probe: FG 16 only)
        a.dbg[(size_t)TM * (nch * NC) + row * 32 + c] = __uint_as_float(gm[c]);
        a.dbg[(size_t)TM * (nch * NC) + row * 32 + 16 + c] = __uint_as_float(gc[c]);
      }
    }
#endif
    const float csum = __uint_as_float(gm[L::ONES_ROW]) + __uint_as_float(gc[L::ONES_ROW]);  // sum cc w
#pragma unroll
    for (int f = 0; f < FM; ++f) {
      const float gsum = __uint_as_float(gm[f]) + __uint_as_float(gc[f]);              // sum cc w s
      const float gex = gacc_w[lane * FG + f - 4 * 32 * FG] + gacc_w[lane * FG + f];  // column halves 0 + 1
      gx[f] = a.rc.grad_scale * (fmaf(xl[f], csum, -gsum) * inv_g + gex);
    }
  }
  tc_fence_before();
  const float score = a.rc.score_scale * ((sc_p[row] + (sc2.lo() + sc2.hi())) * (1.f / 1024.f) +
                                          (sacc_w[lane - 4 * 32] + sacc_w[lane]));
  if (ti + 1 < ntile) named_arrive(TCB_ACC, QT);  // the lower half may reset its accumulators for tile ti + 1
  named_sync(TCB_OWN, TM);  // every owner has read its accumulators: their region becomes `os`
  if (row < nq) {
    float* rec = os + row * n_out;
    rec[0] = score;
    if constexpr (MODE == TC_GRAD) {
      const float scale = (a.grad_out != nullptr) ? a.grad_out[b_base + row] : 1.f;
      if (!has_fk) {
        for (int f = 0; f < F; ++f) rec[1 + f] = scale * gx[f];
      } else {
        float gq[DC_MAX_DOF];
#pragma unroll
        for (int i = 0; i < DC_MAX_DOF; ++i) gq[i] = 0.f;
        if (FG == 16 && a.fk.type == DC_FK_PLANAR_CHAIN && a.fk.n_repeat <= 1 && !a.fk.time_last) {
          fk_planar_vjp_reg<7>(a.fk.n_links, xl, gx, gq);  // the BASELINE robot: J^T in registers
#pragma unroll
          for (int i = 0; i < 7; ++i)
            if (i < a.n_in) rec[1 + i] = scale * gq[i];
        } else {
          float qv[DC_MAX_DOF];
          const float* qs = qs_all + buf * TM * L::QS_DOF;
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qs[row * a.n_in + i] : 0.f;
          fk_vjp<float>(a.fk, qv, xl, 1, gx, 1, gq);
          for (int i = 0; i < a.n_in; ++i) rec[1 + i] = scale * gq[i];
        }
      }
    }
  }
  DC_TC_TRACE_TILE(4, 12);
  named_sync(TCB_OWN, TM);
  const int otid = tid - TM;  // 0 .. 127
  if (fused) {
    // n_bcast > 0: the same block goes to every rank's gathered buffer (peer stores over NVLink) — the all-gather of
    // the multi-GPU path, overlapped with the other tiles' arithmetic
    const int n_dst = a.n_bcast > 0 ? a.n_bcast : 1;
    const size_t off = (size_t)(a.score - (a.n_bcast > 0 ? a.bcast[0] : a.score)) + (size_t)b_base * n_out;
    const int n_words = nq * n_out;
    if (a.n_bcast > 0 && (n_words & 3) == 0 && (off & 3) == 0) {
      // one bulk TMA store of the whole block per destination (shared -> peer global, large NVLink packets, no
      // thread is held by the transfer); the bases are 16-byte aligned (dc_score_grad_bcast), so `off & 3` decides
      if (otid == 0) {
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
          for (int i = otid; i < n_words / 4; i += TM) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(os)[i];
        } else {
          for (int i = otid; i < n_words; i += TM) dst[i] = os[i];
        }
      }
    }
  } else {
    if (otid < nq) a.score[(size_t)(b_base + otid) * a.score_ld] = os[otid * n_out];
    if constexpr (MODE == TC_GRAD) {
      for (int i = otid; i < nq * a.n_in; i += TM) {
        const int tq = i / a.n_in, c = i - tq * a.n_in;
        a.grad[(size_t)(b_base + tq) * a.grad_ld + c] = os[tq * n_out + 1 + c];
      }
    }
  }
  if (a.mirror != nullptr) {
    float* dst = a.mirror + (size_t)b_base * n_out;
    const int n_words = nq * n_out;
    if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int i = otid; i < n_words / 4; i += TM) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(os)[i];
    } else {
      for (int i = otid; i < n_words; i += TM) dst[i] = os[i];
    }
  }
  if (has_fk) mbar_arrive(reinterpret_cast<uint64_t*>(smem + L::SM_BAR) + 15 + buf);  // qs[buf] may be refilled (tile ti + 2)
  named_sync(TCB_OWN, TM);  // `os` is consumed: the owners' accumulators are reset by the next tile
  DC_TC_TRACE_TILE(4, 13);
}

template <int MODE, int FG = 16>
__global__ void __launch_bounds__(TcLayoutT<FG>::THREADS, TcLayoutT<FG>::CTAS_PER_SM) score_tc_kernel(const __grid_constant__ TcArgs a) {
  using L = TcLayoutT<FG>;
  constexpr int NC = L::NC, QT = L::QTHREADS, TM = L::TM;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::SM_BAR);
  uint64_t* bar_b1full = bars;        // [2]   TMA -> GEMM1
  uint64_t* bar_b2full = bars + 2;    // [2]   TMA -> GEMM2
  uint64_t* bar_b2free = bars + 4;    // [2]   GEMM2 done -> TMA
  uint64_t* bar_rho = bars + 6;       // [2]   GEMM1 done -> query threads (and -> TMA: the B1 slot is free)
  uint64_t* bar_cc = bars + 8;        // [2]   query threads -> GEMM2
  uint64_t* bar_a = bars + 10;        // [1]   A operand written -> GEMM1
  uint64_t* bar_g = bars + 11;        // [2]   last GEMM2 of the tile done -> epilogue
  uint64_t* bar_qfull = bars + 13;    // [2]   configurations of a tile staged -> FK stage
  uint64_t* bar_qfree = bars + 15;    // [2]   epilogue done with the staged configurations -> prefetch warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::SM_TMEM_SLOT);
  unsigned char* ring1 = smem + L::SM_RING1;
  unsigned char* ring2 = smem + L::SM_RING2;
  unsigned char* a_op = smem + L::SM_A;
  float* sacc = reinterpret_cast<float*>(smem + L::SM_ACC);      // [8][32]
  float* gacc = sacc + L::QWARPS * 32;                           // [8][32][FG]
  float* sc_p = reinterpret_cast<float*>(smem + L::SM_ROWS);     // [128] score partial of the lower column half
  const float2* thr_all = reinterpret_cast<const float2*>(sc_p + TM);  // [2][128] near-threshold line (c0, c1) of each query
  uint32_t* queues = reinterpret_cast<uint32_t*>(smem + L::SM_QUEUE);

  const long long t0 = (long long)blockIdx.x * a.n_tiles / gridDim.x;
  const long long t1 = (long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x;
  const int ntile = (int)(t1 - t0);
  const int nch = a.n_chunks;
#ifdef DC_TC_ENABLE_TRACE
  if (a.trace != nullptr && tid == 0) {  // per-CTA residency record: SM id, start / end of the CTA (globaltimer, ns)
    unsigned smid;
    unsigned long long now;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    a.trace[2048 + blockIdx.x * 4 + 0] = smid;
    a.trace[2048 + blockIdx.x * 4 + 1] = (long long)now;
    a.trace[2048 + blockIdx.x * 4 + 3] = ntile;
  }
#endif
  const uint32_t total = (uint32_t)ntile * (uint32_t)nch;  // < 2^31 (checked by launch_score_tc)
  const unsigned char* blob1 = a.blob;
  const unsigned char* blob2 = a.blob + (size_t)nch * L::B1_BYTES;

  if (warp == L::CTRL_WARP) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_b1full[i], 1);
        mbar_init(&bar_b2full[i], 1);
        mbar_init(&bar_b2free[i], 1);
        mbar_init(&bar_rho[i], 1);
        mbar_init(&bar_cc[i], QT);
        mbar_init(&bar_g[i], 1);
        mbar_init(&bar_qfull[i], 32);
        mbar_init(&bar_qfree[i], TM);
      }
      mbar_init(bar_a, TM);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, L::TMEM_COLS);
  } else if (warp == 0 && lane < 12) {
    reinterpret_cast<float*>(smem + L::SM_TRAILER)[lane] = tc_trailer(a.blob, a.n_sv, FG)[lane];
  } else if (warp == 1 && lane == 0) {
    *reinterpret_cast<long long*>(smem + L::SM_TILE0) = t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Register reallocation between the warpgroups (setmaxnreg): the kernel is launched with 80 registers per thread
  // (12 warps x 2 CTAs per SM); the service warpgroup keeps 32 and the two query warpgroups grow to 104.
  if (warp >= L::CTRL_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    if (warp == L::CTRL_WARP) {
      // ================= MMA issuer: GEMM2(j) then GEMM1(j + 2), in that order, from ONE thread ======================
      constexpr uint32_t idesc1 = umma_idesc_f16(NC);
      constexpr uint32_t idesc2 = umma_idesc_f16(L::N2);
      const uint32_t a_s = smem_u32(a_op);
      auto gemm1 = [&](uint32_t gg) {
        const int st = (int)(gg & 1);
        mbar_wait_wd(&bar_b1full[st], (uint32_t)((gg >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_s = smem_u32(ring1 + (size_t)st * L::B1_BYTES);
          const uint32_t d = tmem + st * L::COL_STAGE;
#pragma unroll
          for (int ks = 0; ks < L::K1 / 16; ++ks) {
            const uint64_t ad = umma_desc(a_s + ks * 2 * (TM * 16), TM * 16, 128);
            const uint64_t bd = umma_desc(b_s + ks * 2 * (NC * 16), NC * 16, 128);
            umma_f16_ss(d, ad, bd, idesc1, ks == 0 ? 0u : 1u);
          }
          umma_commit(&bar_rho[st]);
        }
        __syncwarp();
      };
      uint32_t g = 0;
      for (int ti = 0; ti < ntile; ++ti) {
        mbar_wait_wd(bar_a, (uint32_t)(ti & 1));
        tc_fence_after();
        gemm1(g);
        if (nch > 1) gemm1(g + 1);
        for (int j = 0; j < nch; ++j, ++g) {
          const int st = (int)(g & 1);
          mbar_wait_wd(&bar_b2full[st], (uint32_t)((g >> 1) & 1));
          mbar_wait_wd(&bar_cc[st], (uint32_t)((g >> 1) & 1));
          tc_fence_after();
          DC_TC_TRACE(3, g);
          if (elect_one()) {
            if constexpr (MODE == TC_GRAD) {
              const uint32_t b_s = smem_u32(ring2 + (size_t)st * L::B2W_BYTES);
              const uint32_t cc = tmem + st * L::COL_STAGE;
              const uint32_t d = tmem + L::COL_G + (uint32_t)((L::G_BUFS == 2) ? (ti & 1) : 0) * L::N2;
#pragma unroll
              for (int ks = 0; ks < L::KS2; ++ks) {
                const uint64_t bd = umma_desc(b_s + ks * L::B2_STEP_BYTES, L::N2 * 16, 128);
                umma_f16_ts(d, cc + ks * 8, bd, idesc2, (j == 0 && ks == 0) ? 0u : 1u);
              }
            }
            umma_commit(&bar_b2free[st]);
            if (j == nch - 1) umma_commit(&bar_g[(L::G_BUFS == 2) ? (ti & 1) : 0]);
          }
          __syncwarp();
          DC_TC_TRACE(4, g);
          if (j + 2 < nch) gemm1(g + 2);
          DC_TC_TRACE(1, g);
        }
      }
    } else if (warp == L::CTRL_WARP + 1) {
      // ================= GEMM1 operand images: slot gg & 1 is free again once GEMM1(gg - 2) has completed ==============
      for (uint32_t gg = 0; gg < total; ++gg) {
        const int st = (int)(gg & 1);
        if (gg >= 2) mbar_wait_wd(&bar_rho[st], (uint32_t)(((gg - 2) >> 1) & 1));
        if (elect_one()) {
          mbar_expect_tx(&bar_b1full[st], L::B1_BYTES);
          tma_bulk_g2s(ring1 + (size_t)st * L::B1_BYTES, blob1 + (size_t)(gg % (uint32_t)nch) * L::B1_BYTES, L::B1_BYTES,
                       &bar_b1full[st]);
        }
        __syncwarp();
      }
    } else if (warp == L::CTRL_WARP + 3) {
      // ================= configurations of tile ti -> shared memory (coalesced 16-byte loads), up to two tiles ahead: the
      // global / PCIe latency (q may be a pinned host buffer, read zero-copy) hides behind the scoring of the earlier tiles
      if (a.fk.type != DC_FK_NONE) {
        float* qs_all = reinterpret_cast<float*>(smem + L::SM_QS);
        for (int ti = 0; ti < ntile; ++ti) {
          const int buf = ti & 1;
          if (ti >= 2) mbar_wait_wd(&bar_qfree[buf], (uint32_t)(((ti - 2) >> 1) & 1));
          const long long b_base = (t0 + ti) * TM;
          const int nq = (int)min((long long)TM, a.batch - b_base);
          const float* src = a.q + (size_t)b_base * a.n_in;
          float* qs = qs_all + buf * TM * L::QS_DOF;
          const int n_words = nq * a.n_in;
          if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            for (int i = lane; i < n_words / 4; i += 32) reinterpret_cast<float4*>(qs)[i] = s4[i];
          } else {
            for (int i = lane; i < n_words; i += 32) qs[i] = src[i];
          }
          mbar_arrive(&bar_qfull[buf]);  // release: the stores above are visible to whoever observes the phase
        }
      }
    } else if (warp == L::CTRL_WARP + 2) {
      // ================= GEMM2 operand images + weights: slot gg & 1 is free again once GEMM2(gg - 2) has completed (the
      // query warps read the weights of chunk gg - 2 before they arrive on bar_cc, which GEMM2(gg - 2) waits for) =============
      for (uint32_t gg = 0; gg < total; ++gg) {
        const int st = (int)(gg & 1);
        if (gg >= 2) mbar_wait_wd(&bar_b2free[st], (uint32_t)(((gg - 2) >> 1) & 1));
        if (elect_one()) {
          mbar_expect_tx(&bar_b2full[st], L::B2W_BYTES);
          tma_bulk_g2s(ring2 + (size_t)st * L::B2W_BYTES, blob2 + (size_t)(gg % (uint32_t)nch) * L::B2W_BYTES, L::B2W_BYTES,
                       &bar_b2full[st]);
        }
        __syncwarp();
      }
    }
  } else {
#ifndef DC_TC_QREGS
#define DC_TC_QREGS 104  // 112 would fill the register file exactly (2 x (8 x 32 x 112 + 4 x 32 x 32) = 65536) and was tried
#endif                   // to stop ptxas rematerialising addresses in the chunk loop: the second CTA's setmaxnreg.inc never
                         // returns (the launch-time allocation of 80 x 384 leaves no slack) — the kernel hangs.  Do not raise.
#define DC_TC_STR2(x) #x
#define DC_TC_STR(x) DC_TC_STR2(x)
    if constexpr (FG == 16) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 " DC_TC_STR(DC_TC_QREGS) ";" ::: "memory");
    } else {  // one CTA per SM: 8 x 32 x 168 + 4 x 32 x 32 = 47 K of the 64 K registers — room to spare, no spills
      asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
    }
    // ================= query threads: row = 32 (warp & 3) + lane, column half = warp >> 2 ==========================
    const int row = ((warp & 3) << 5) | lane;
    const int hcol = warp >> 2;
    // owners (upper half) run the epilogue of their row; the lower half runs FK + the A operand of the NEXT tile meanwhile
    const bool owner = hcol == 1;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t* queue = queues + warp * L::QCAP;
    float* gacc_w = gacc + warp * 32 * FG;
    float* sacc_w = sacc + warp * 32;
#ifdef DC_TC_ENABLE_TRACE
    const float* trailer = tc_trailer(a.blob, a.n_sv, FG);
    const float tau = trailer[5], inv_tc0 = trailer[9];
#endif

    if (!owner && ntile > 0) tc_fk_stage<FG>(a, 0, tid);

    uint32_t g = 0;
    for (int ti = 0; ti < ntile; ++ti) {
      const int buf = ti & 1;
      if (owner) {
        named_sync(TCB_FK, QT);
      } else if (ti > 0) {
        named_sync(TCB_ACC, QT);  // the owners have read the lower half's exact accumulators of tile ti - 1
      }
      DC_TC_TRACE_TILE(0, 0);
      DC_TC_TRACE_TILE(4, 8);
      {
        float4* z = reinterpret_cast<float4*>(gacc_w + lane * FG);
#pragma unroll
        for (int v = 0; v < FG / 4; ++v) z[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        sacc_w[lane] = 0.f;
      }
      __syncwarp();
#if DC_TC_DEPHASE
      // The two halves run half a chunk apart: when one half sits in the per-chunk waits (mbarrier, tcgen05.ld / st round
      // trips) the other is in its arithmetic, instead of both stalling the SM sub-partition at the same moments.
      if (owner) named_sync(TCB_PHASE, QT);
#endif

      P2 sc2(0.f, 0.f);
      int qcount = 0;
      // two register buffers of 16 columns: the load of batch b + 1 is in flight while batch b is processed
      uint32_t ra[16], rc[16];
      for (int j = 0; j < nch; ++j, ++g) {
        const int st = (int)(g & 1);
        // this chunk's weights (warp-uniform LDS.128 broadcasts) and max|s|^2 ride with the GEMM2 image
        const float* wsm = reinterpret_cast<const float*>(ring2 + (size_t)st * L::B2W_BYTES + L::OFF_W);
        const float4* wv4 = reinterpret_cast<const float4*>(wsm + hcol * (NC / 2));
        if (warp == 0) DC_TC_TRACE(8, g);
        mbar_wait_wd(&bar_b2full[st], (uint32_t)((g >> 1) & 1));
        const float s2j = wsm[L::META_S2MAX];
        const uint32_t tcol = tm_lane + st * L::COL_STAGE + hcol * (NC / 2);
        if (!DC_TC_PREFETCH || j == 0) {  // (otherwise issued at the end of the previous chunk)
          mbar_wait_wd(&bar_rho[st], (uint32_t)((g >> 1) & 1));
          tc_fence_after();
          tmem_ld16(tcol, ra);
          tmem_ld16(tcol + 16, rc);
        }
        if (warp == 0) DC_TC_TRACE(10, g);
        // near threshold of this (query, chunk) on T (the line was set up by tc_fk_stage; kept in shared memory, not in
        // registers: the chunk loop has none to spare, and a spilled register would compete with the weights for the L1)
        const float2 tl = thr_all[buf * TM + row];
        const float thr = fmaf(tl.y, s2j, tl.x);
        auto batch = [&](uint32_t* rb, const int bt) {
          const int col0 = bt * 16;
#ifdef DC_TC_ENABLE_TRACE
          if (a.dbg != nullptr && t0 + ti == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              a.dbg[(size_t)row * (nch * NC) + j * NC + hcol * (NC / 2) + col0 + c] = (__uint_as_float(rb[c]) - tau) * inv_tc0;
          }
#endif
          // ---- near pairs: queued for exact evaluation, removed from the tensor-core sums (T := huge -> u = 0) ----
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
              if (near) {
                queue[qcount + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)row << 24) | (uint32_t)(n0 + c);
              }
              if (lane == 0) {  // the rows are read when the queue is drained: start pulling them into L1 now
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.table + (size_t)(n0 + c) * a.row_stride));
                if (a.table_lo != nullptr)
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(a.table_lo + (size_t)(n0 + c) * a.row_stride));
              }
              qcount += __popc(bal);
              if (qcount > L::QCAP - 32) {
                __syncwarp();
                tc_drain_pairs<FG>(a, qcount, warp, buf);
                __syncwarp();
                qcount = 0;
              }
            }
          }
          // ---- all pairs: radial profile on the tensor-core T, packed over column pairs ------------------------
          uint32_t outp[16];  // per 8 supports: 4 x f16x2 ch, 4 x f16x2 cl  (GEMM2 K slots 0..7, 8..15)
          float4 w4[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) w4[c] = wv4[bt * 4 + c];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 w2 = (c & 2) ? make_float2(w4[c >> 2].z, w4[c >> 2].w) : make_float2(w4[c >> 2].x, w4[c >> 2].y);
            const float ta = __uint_as_float(rb[c]), tb = __uint_as_float(rb[c + 1]);
            const float rr = fast_rcp(ta * tb);  // one MUFU per column pair: 1/ta = tb / (ta tb)
            const P2 u = pmul_b(P2(tb, ta), rr);  // 32 u
            const P2 k = pmul(u, u);              // 2^10 u^2
            sc2 = pfma(P2(w2), k, sc2);
            if constexpr (MODE == TC_GRAD) {
              // cc = 2^15 u^3 split into two 11-bit f16 terms WITHOUT the conversion unit (F2FP shares the XU pipe with
              // MUFU and was the kernel's binding pipe, profiles/r02e_*): scaling by 2^-112 re-biases the fp32 exponent
              // to f16's, after which `bits >> 13` IS the f16 bit pattern (truncated; f16 denormals included).
              const int o = (c >> 3) * 8 + ((c & 7) >> 1);
#if DC_TC_INTPACK
              const P2 cc = pmul_b(pmul(k, u), 0x1p-112f);
              const uint32_t b0 = __float_as_uint(cc.lo()) & 0xffffe000u, b1 = __float_as_uint(cc.hi()) & 0xffffe000u;
              const P2 cl = padd(cc, P2(-__uint_as_float(b0), -__uint_as_float(b1)));  // exact
              outp[o] = b1 * 8u + (b0 >> 13);  // {ch(c + 1) : ch(c)} as f16x2 (the low 13 bits of b1 are zero)
              outp[o + 4] = __byte_perm(__float_as_uint(cl.lo()) << 3, __float_as_uint(cl.hi()) << 3, 0x7632);
#else
              const P2 cc = pmul(k, u);           // 2^15 u^3
              const P2 ch(split_hi(cc.lo()), split_hi(cc.hi()));
              const P2 cl = padd(cc, P2(-ch.lo(), -ch.hi()));
              outp[o] = pack_f16x2(ch.lo(), ch.hi());
              outp[o + 4] = pack_f16x2(cl.lo(), cl.hi());
#endif
            }
          }
          if constexpr (MODE == TC_GRAD) tmem_st16(tcol + col0, outp);
        };
        tmem_wait_ld();
        if (warp == 0) DC_TC_TRACE(11, g);
        batch(ra, 0);
        tmem_ld16(tcol + 32, ra);  // batch 2 into the first buffer (its values are consumed)
        batch(rc, 1);
#if DC_TC_DEPHASE
        if (j == 0 && !owner) named_arrive(TCB_PHASE, QT);  // the owners start their chunk 0 now
#endif
        tmem_wait_ld();
        batch(ra, 2);
#if DC_TC_PREFETCH
        if (j + 1 < nch) {
          // T of chunk g + 1 (other stage) was issued two chunks ago and is normally complete: start pulling its first 32
          // columns now; the loads land while this chunk's stores drain and the arrive / loop top execute
          mbar_wait_wd(&bar_rho[st ^ 1], (uint32_t)(((g + 1) >> 1) & 1));
          tc_fence_after();
          const uint32_t tnext = tm_lane + (st ^ 1) * L::COL_STAGE + hcol * (NC / 2);
          tmem_ld16(tnext, ra);
          tmem_ld16(tnext + 16, rc);
        }
#endif
        if (warp == 0) DC_TC_TRACE(12, g);
        if constexpr (MODE == TC_GRAD) tmem_wait_st();
        tc_fence_before();
        if (warp == 0) DC_TC_TRACE(13, g);
        mbar_arrive(&bar_cc[st]);
      }
      DC_TC_TRACE_TILE(0, 1);
      DC_TC_TRACE_TILE(4, 9);
      // (DC_TC_HANDOVER, measured, off: with another tile to come the lower half hands what is left in its queue to the
      // owner warp of the same TMEM lane quarter and goes straight to the FK stage; that warp applies the entries to the
      // LOWER warp's accumulators, in queue order, before it reads them — same additions, same order, same bits.)
      const bool hand_over = DC_TC_HANDOVER && !owner && (ti + 1 < ntile);
      if (qcount > 0 && !hand_over) {
        __syncwarp();
        tc_drain_pairs<FG>(a, qcount, warp, buf);
      }
      DC_TC_TRACE_TILE(0, 2);
      DC_TC_TRACE_TILE(4, 10);

      if (!owner) {
        if (lane == 0) reinterpret_cast<int*>(smem + L::SM_QHAND)[warp] = hand_over ? qcount : 0;
        sc_p[row] = sc2.lo() + sc2.hi();
        // bar.arrive orders this thread's prior shared-memory writes for the threads that complete the barrier; a
        // sequentially consistent fence here (MEMBAR.SC) cost ~2000 cycles per tile on the path to the next tile's FK
        named_arrive(TCB_EPI, QT);  // lower half's partial scores and exact terms of tile ti are complete
        if (ti + 1 < ntile) tc_fk_stage<FG>(a, ti + 1, tid);
        DC_TC_TRACE_TILE(0, 3);
        continue;
      }

      tc_epilogue<MODE, FG>(a, ti, ntile, sc2.lo() + sc2.hi());
    }
    if (tid == TM && a.n_bcast > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // peer stores performed
    if (a.sync_world > 0 && owner) {
      // every record this CTA sent (bulk stores of thread TM, per-thread stores of all owners on the unaligned path) is
      // ordered before the count; the last CTA of the grid then runs the flag exchange for the whole launch
      __threadfence_system();
      named_sync(TCB_OWN, TM);
      if (tid == TM) {
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {
          *a.done = 0u;  // for the next launch (stream order)
          __threadfence_system();
          for (int r = 0; r < a.sync_world; ++r)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag_peer[r]), "r"(a.epoch) : "memory");
          const long long t_start = clock64();
          for (int r = 0; r < a.sync_world; ++r) {
            for (;;) {
              uint32_t v;
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.flag_mine + r) : "memory");
              if ((int32_t)(v - a.epoch) >= 0) break;
              if (clock64() - t_start > a.sync_timeout_cycles) __trap();  // a rank that never arrives must not hang the GPU
            }
          }
        }
      }
    }
  }

  // teardown: the MMA issuer (owner of the TMEM allocation) waits for the query warps; the TMA warps just leave
  if (warp <= L::CTRL_WARP) {
    tc_fence_before();
    named_sync(TCB_END, QT + 32);
#ifdef DC_TC_ENABLE_TRACE
    if (a.trace != nullptr && tid == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      a.trace[2048 + blockIdx.x * 4 + 2] = (long long)now;
    }
#endif
    if (warp == L::CTRL_WARP) {
      tc_fence_after();
      tmem_dealloc(tmem, L::TMEM_COLS);
    }
  }
}

// Packs S_feat[N,F], w[N] for RQKernel(gamma, p = 2) into `blob` (tc_blob_bytes(N) bytes, 128-byte aligned).
inline int launch_pack_supports_tc(const float* s_feat, const float* w, long long n, int F, float gamma, unsigned char* blob,
                                   cudaStream_t stream) {
  using L = TcLayout;
  const int nch = tc_n_chunks(n), fg = tc_group(F);
  DC_CUDA_OK(cudaMemsetAsync(blob, 0, tc_blob_bytes(n, fg), stream));
  int* trailer = reinterpret_cast<int*>(const_cast<float*>(tc_trailer(blob, n, fg)));
  tc_scan_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(s_feat, w, (int)n, F, trailer);
  DC_LAUNCH_CHECK();
  const unsigned grid = (unsigned)((nch * L::NC + 127) / 128);
  if (fg == 16)
    pack_supports_tc_kernel<16><<<grid, 128, 0, stream>>>(s_feat, w, (int)n, F, nch, gamma, blob);
  else
    pack_supports_tc_kernel<32><<<grid, 128, 0, stream>>>(s_feat, w, (int)n, F, nch, gamma, blob);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

template <int MODE, int FG>
int launch_score_tc_fg(TcArgs& a, int num_sms, cudaStream_t stream) {
  using L = TcLayoutT<FG>;
  a.n_tiles = (int)ceil_div64(a.batch, L::TM);
  a.n_chunks = tc_n_chunks(a.n_sv);
  const int grid = (int)min((long long)L::CTAS_PER_SM * num_sms, (long long)a.n_tiles);
  if (((long long)a.n_tiles / grid + 1) * a.n_chunks >= (1LL << 31) || a.n_sv >= (1 << 24)) return DC_ERR_UNSUPPORTED;
  if (a.n_feat > L::FMAX || (a.fk.type != DC_FK_NONE && a.n_in > L::QS_DOF)) return DC_ERR_UNSUPPORTED;
  auto kern = score_tc_kernel<MODE, FG>;
  {
    static PerDeviceOnce once;
    int dev = 0;
    const int p = once.pending(&dev);
    if (p < 0) return DC_ERR_NO_DEVICE;
    if (p > 0) {
      DC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SM_BYTES));
      DC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      once.mark(dev);
    }
  }
  kern<<<grid, L::THREADS, L::SM_BYTES, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// F <= 14: operand groups of 16 K slots, two CTAs per SM (the BASELINE shape); F <= 30: groups of 32, one CTA per SM.
template <int MODE>
int launch_score_tc(TcArgs& a, int num_sms, cudaStream_t stream) {
  return tc_group(a.n_feat) == 16 ? launch_score_tc_fg<MODE, 16>(a, num_sms, stream)
                                  : launch_score_tc_fg<MODE, 32>(a, num_sms, stream);
}

}  // namespace dc
