// Fused collision-score kernel, thread-per-query form (large batches, fp32).
//
//   score[b,c] = sum_n w[n,c] k(|FK(q_b) - s_n|^2)         grad[b,:] = J_FK(q_b)^T sum_n omega_n coef_n (x_b - s_n)
//
// replaces, in ONE launch and without ever materialising the B x N kernel matrix, what the reference does with
// fkine -> torch.cdist -> pow/reciprocal -> matmul and the autograd backward of all of them
// (diffco/kernel_perceptrons.py:309-319,362-370; diffco/kernel.py:17-79; diffco/model.py:40-48,225-241).
//
// Work decomposition (B200: 148 SMs, one persistent 16-warp CTA per SM):
//   * the batch is cut into tiles of QT = 64 queries; CTA i owns a contiguous, balanced range of tiles and deals them
//     round-robin to its 16/NWG warp groups; a group synchronises only with itself (named barriers);
//   * per tile, phase A runs FK one-query-per-thread into shared memory (xs[F][64], conflict-free columns) and phase C
//     runs the J^T product the same way;
//   * phase B: lane l of every warp of the group holds queries (2l, 2l+1) of the tile as the two halves of packed
//     registers, and the group's NWG warps split the support set into NWG contiguous slices;
//   * each warp streams its own slice of the packed support table HBM/L2 -> shared memory with 1-D bulk TMA
//     (cp.async.bulk + mbarrier complete_tx) through a private STAGES-deep ring, running STAGES chunks ahead and
//     prefetching across tile boundaries; rows are read back as warp-uniform LDS.128 broadcasts;
//   * the pair update is 2-wide packed FP32 throughout (FADD2 / FMUL2 / FFMA2 with the support value or weight as
//     the broadcast scalar operand): per support vector and query pair F subtracts, F fused squares, ~6 packed ops
//     for the radial profile and the score, 2 MUFU, F fused gradient accumulations — no scalar FMA-pipe
//     instructions in the loop (mixing them with packed ones costs ~20% of the pipe, tools/ubench/fma_pipes.cu);
//   * partials are combined across warps through shared memory in a fixed order (deterministic results, and a row's
//     result does not depend on its position in the batch).
#pragma once

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

namespace dc {

enum ScoreMode { M_SCORE = 0, M_GRAD = 1 };

template <typename T>
struct ScoreArgs {
  dc_fk_desc fk;
  RadialConsts<T> rc;
  const T* table;
  const T* q;
  T* score;
  T* grad;
  const T* grad_out;
  long long batch;
  long long score_ld;  // elements between consecutive rows of score (>= C)
  long long grad_ld;   // elements between consecutive rows of grad  (>= D, or >= C*D in Jacobian mode)
  int n_sv;
  int n_feat;     // F
  int n_class;    // C
  int n_in;       // columns of q (dof, or F when fk.type == NONE)
  int jac_class;  // >= 0: gradient of class jac_class alone, written at grad + b*grad_ld + jac_class*n_in (Jacobian pass)
  int write_score;
  int n_tiles;    // ceil(batch / QT)
  int st_q;       // super-tile size in queries (multiple of QT, <= NT)
  int chunk_rows;
};

// NWG = warps per group (4 or NW).  A CTA has NW = 16 warps (8 for F > 16: the per-lane state of 3F packed registers then
// needs the 255 registers a 256-thread CTA leaves per thread) = NW/NWG groups; a group owns one 64-query tile at a time
// and its NWG warps split the support set.  NWG = 4 (four concurrent tiles per SM, group-local named barriers, 4-way
// reduction) is the throughput configuration; NWG = 16 (one tile per SM at a time) keeps every warp busy when the
// batch has fewer than ~4 tiles per SM.
template <int F, int CW, int MODE, int NWG, int STAGES>
struct TqCfg {
  static constexpr int NW = F > 16 ? 8 : 16;
  static constexpr int NGRP = NW / NWG;
  static constexpr int GT = NWG * 32;  // threads per group
  static constexpr int FPAD = round_up(F, 2);
  static constexpr int ROW = round_up(FPAD + CW, 4);
  static constexpr int QT = 64;
  static constexpr int NG = (MODE == M_GRAD) ? 1 : 0;
  static constexpr int NRED = CW + NG * F;
  static constexpr int BAR_BYTES = round_up(NW * STAGES * 8, 128);
  static constexpr int QS = QT * DC_MAX_DOF;      // staged configurations of the tile (per group)
  static constexpr int GS = (CW + NG * F) * QT;   // reduced scores + feature gradients of the tile (per group)
  // per-group scratch: NWG partial sums per reduced value, later re-used for the tile's output records [QT][C + D]
  static constexpr int RED = QT * (NWG * NRED > CW + NG * DC_MAX_DOF ? NWG * NRED : CW + NG * DC_MAX_DOF);
  __host__ __device__ static constexpr size_t smem_bytes(int chunk_rows) {
    return (size_t)BAR_BYTES + sizeof(float) * ((size_t)NW * STAGES * chunk_rows * ROW + (size_t)NGRP * (RED + F * QT + GS + QS));
  }
};

__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int F, int KIND, int CW, int MODE, int NWG, int STAGES>
__global__ void __launch_bounds__(TqCfg<F, CW, MODE, NWG, STAGES>::NW * 32, 1) score_tq_kernel(const __grid_constant__ ScoreArgs<float> a) {
  using T = float;
  using Cfg = TqCfg<F, CW, MODE, NWG, STAGES>;
  constexpr int NW = Cfg::NW, NGRP = Cfg::NGRP, GT = Cfg::GT;
  constexpr int FPAD = Cfg::FPAD, ROW = Cfg::ROW, QT = Cfg::QT, NG = Cfg::NG, NRED = Cfg::NRED;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp / NWG, wig = warp - grp * NWG, gtid = tid - grp * GT;  // group, warp in group, thread in group
  const int CH = a.chunk_rows;
  const int bar_id = 1 + grp;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * STAGES;
  T* ring_all = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES);
  T* ring = ring_all + (size_t)warp * STAGES * CH * ROW;
  T* red = ring_all + (size_t)NW * STAGES * CH * ROW + (size_t)grp * Cfg::RED;  // [wig][kk][QT]
  T* os = red;                                                                          // [QT][rec] after the reduction
  T* xs = ring_all + (size_t)NW * STAGES * CH * ROW + (size_t)NGRP * Cfg::RED + (size_t)grp * (F * QT + Cfg::GS + Cfg::QS);
  T* gs = xs + (size_t)F * QT;  // [CW + NG*F][QT]: reduced scores, then feature gradients
  T* qs = gs + Cfg::GS;         // [QT][n_in]
  // Output addressing.  `fused`: score and grad are the two halves of one [B, C+D] record buffer, so a tile's output is
  // ONE contiguous block (what the all-gather ships, and what makes zero-copy writes to pinned host memory efficient).
  const int n_out = a.n_class + (NG ? a.n_in : 0);
  const bool fused = (NG == 0) ? (a.score_ld == a.n_class)
                               : (a.jac_class < 0 && a.score_ld == a.grad_ld && a.score_ld == n_out && a.grad == a.score + a.n_class);

  // ---- this CTA's tiles (contiguous, balanced), dealt round-robin to its groups; this warp's slice of the supports
  const long long t0 = (long long)blockIdx.x * a.n_tiles / gridDim.x;
  const long long t1 = (long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x;
  const long long my_tiles = (t1 - t0 > grp) ? (t1 - t0 - grp + NGRP - 1) / NGRP : 0;
  const int n0 = (int)((long long)wig * a.n_sv / NWG);
  const int n1 = (int)((long long)(wig + 1) * a.n_sv / NWG);
  const int n_chunks = (n1 - n0 + CH - 1) / CH;
  const long long total_chunks = (long long)n_chunks * my_tiles;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  // producer state (lane 0): next chunk of the slice to fetch and the stage it goes to; chunks are fetched in slice
  // order over and over, once per tile, so the ring runs ahead across tile boundaries
  long long issued = 0;
  int ci_issue = 0, st_issue = 0;
  auto issue = [&]() {
    const int r0 = n0 + ci_issue * CH;
    const int rows = min(CH, n1 - r0);
    const uint32_t bytes = (uint32_t)rows * ROW * sizeof(T);
    mbar_expect_tx(&bars[st_issue], bytes);
    tma_bulk_g2s(ring + (size_t)st_issue * CH * ROW, a.table + (size_t)r0 * ROW, bytes, &bars[st_issue]);
    ++issued;
    if (++ci_issue == n_chunks) ci_issue = 0;
    if (++st_issue == STAGES) st_issue = 0;
  };
  if (lane == 0) {
    for (int s = 0; s < STAGES && issued < total_chunks; ++s) issue();
  }
  int st_use = 0;
  uint32_t parity = 0;

  for (long long tile = t0 + grp; tile < t1; tile += NGRP) {
    const long long b_base = tile * QT;
    const int nq = (int)min((long long)QT, a.batch - b_base);

    // ---- phase A: stage the tile's configurations (coalesced; device or mapped host memory), then FK, one query per
    //      thread, features into xs[f][t] ----------------------------------------------------------------------------
    {
      const T* src = a.q + (size_t)b_base * a.n_in;
      const int n_words = nq * a.n_in;
      if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = gtid; i < n_words / 4; i += GT) reinterpret_cast<float4*>(qs)[i] = s4[i];
      } else {
        for (int i = gtid; i < n_words; i += GT) qs[i] = src[i];
      }
    }
    group_barrier(bar_id, GT);
    if (gtid < QT) {
      if (gtid < nq) {
        const T* qp = qs + gtid * a.n_in;
        if (a.fk.type == DC_FK_NONE) {
#pragma unroll
          for (int f = 0; f < F; ++f) xs[f * QT + gtid] = qp[f];
        } else {
          T qv[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qp[i] : (T)0;
          fk_features(a.fk, qv, xs + gtid, QT);
        }
      } else {
#pragma unroll
        for (int f = 0; f < F; ++f) xs[f * QT + gtid] = (T)0;
      }
    }
    group_barrier(bar_id, GT);

    // ---- phase B: this warp's slice of the supports against the tile's 64 queries (2 per lane) ------
    P2 x[F];
    P2 sc[CW];
    P2 g[NG > 0 ? F : 1];
    P2 go[CW];
#pragma unroll
    for (int f = 0; f < F; ++f) x[f] = P2(*reinterpret_cast<const float2*>(xs + f * QT + 2 * lane));
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      sc[c] = P2(0.f, 0.f);
      go[c] = P2(1.f, 1.f);
    }
#pragma unroll
    for (int f = 0; f < (NG > 0 ? F : 1); ++f) g[f] = P2(0.f, 0.f);
    if constexpr (MODE == M_GRAD && CW > 1) {
      if (a.jac_class >= 0) {
#pragma unroll
        for (int c = 0; c < CW; ++c) go[c] = (c == a.jac_class) ? P2(1.f, 1.f) : P2(0.f, 0.f);
      } else {
        const long long b = b_base + 2 * lane;
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          T v0 = (c < a.n_class) ? (T)1 : (T)0, v1 = v0;
          if (a.grad_out != nullptr && c < a.n_class) {
            v0 = (b < a.batch) ? a.grad_out[(size_t)b * a.n_class + c] : (T)0;
            v1 = (b + 1 < a.batch) ? a.grad_out[(size_t)(b + 1) * a.n_class + c] : (T)0;
          }
          go[c] = P2(v0, v1);
        }
      }
    }

    for (int ci = 0; ci < n_chunks; ++ci) {
      mbar_wait(&bars[st_use], parity);
      const T* buf = ring + (size_t)st_use * CH * ROW;
      const int rows = min(CH, n1 - (n0 + ci * CH));
#pragma unroll 2
      for (int r = 0; r < rows; ++r) {
        T rowv[ROW];
        const float4* rp = reinterpret_cast<const float4*>(buf + (size_t)r * ROW);
#pragma unroll
        for (int i = 0; i < ROW / 4; ++i) {
          const float4 v = rp[i];
          rowv[4 * i] = v.x;
          rowv[4 * i + 1] = v.y;
          rowv[4 * i + 2] = v.z;
          rowv[4 * i + 3] = v.w;
        }
        P2 d[F];
#pragma unroll
        for (int f = 0; f < F; ++f) d[f] = padd_b(x[f], rowv[f]);  // the table holds -s
        P2 acc0 = pmul(d[0], d[0]);
        P2 acc1(0.f, 0.f);
#pragma unroll
        for (int f = 1; f < F; ++f) {
          if (f & 1)
            acc1 = (f == 1) ? pmul(d[f], d[f]) : pfma(d[f], d[f], acc1);
          else
            acc0 = pfma(d[f], d[f], acc0);
        }
        if constexpr (F > 1) acc0 = padd(acc0, acc1);
        P2 k, coef;
        radial_eval2<KIND>(a.rc, acc0, k, coef);
        if constexpr (CW == 1) {
          const T w = rowv[FPAD];
          sc[0] = pfma_b(w, k, sc[0]);
          if constexpr (MODE == M_GRAD) {
            const P2 cc = pmul_b(coef, w);
#pragma unroll
            for (int f = 0; f < F; ++f) g[f] = pfma(cc, d[f], g[f]);
          }
        } else {
          P2 om(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const T w = rowv[FPAD + c];
            sc[c] = pfma_b(w, k, sc[c]);
            if constexpr (MODE == M_GRAD) om = pfma_b(w, go[c], om);
          }
          if constexpr (MODE == M_GRAD) {
            const P2 cc = pmul(om, coef);
#pragma unroll
            for (int f = 0; f < F; ++f) g[f] = pfma(cc, d[f], g[f]);
          }
        }
      }
      __syncwarp();
      if (lane == 0 && issued < total_chunks) {
        fence_proxy_async();
        issue();  // refills exactly the stage that was just drained
      }
      if (++st_use == STAGES) {
        st_use = 0;
        parity ^= 1u;
      }
    }

    // ---- combine the NWG partial sums of every query (fixed order) -----------------------------------
    {
      float2* rq = reinterpret_cast<float2*>(red + (size_t)wig * NRED * QT) + lane;
#pragma unroll
      for (int c = 0; c < CW; ++c) rq[c * (QT / 2)] = sc[c].v;
#pragma unroll
      for (int f = 0; f < NG * F; ++f) rq[(CW + f) * (QT / 2)] = g[f].v;
    }
    group_barrier(bar_id, GT);
    for (int idx = gtid; idx < NRED * QT; idx += GT) {
      const int kk = idx / QT, qi = idx - kk * QT;
      T s = (T)0;
#pragma unroll
      for (int w = 0; w < NWG; ++w) s += red[(w * NRED + kk) * QT + qi];
      gs[idx] = s * (kk < CW ? a.rc.score_scale : a.rc.grad_scale);
    }
    group_barrier(bar_id, GT);  // `red` is dead from here on: its storage becomes the output staging block `os`

    // ---- phase C: J_FK^T g_x, one query per thread, records into os[t][score | grad] ----------------------
    if (gtid < nq) {
      T* rec = os + gtid * n_out;
      for (int c = 0; c < a.n_class; ++c) rec[c] = gs[c * QT + gtid];
      if constexpr (MODE == M_GRAD) {
        const T* qp = qs + gtid * a.n_in;
        T scale = (T)1;
        if (CW == 1 && a.grad_out != nullptr) scale = a.grad_out[b_base + gtid];
        T* out = rec + a.n_class;
        if (a.fk.type == DC_FK_NONE) {
#pragma unroll
          for (int f = 0; f < F; ++f) out[f] = scale * gs[(CW + f) * QT + gtid];
        } else {
          T qv[DC_MAX_DOF], gq[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) {
            qv[i] = (i < a.n_in) ? qp[i] : (T)0;
            gq[i] = (T)0;
          }
          fk_vjp<T>(a.fk, qv, xs + gtid, QT, gs + CW * QT + gtid, QT, gq);
          for (int i = 0; i < a.n_in; ++i) out[i] = scale * gq[i];
        }
      }
    }
    group_barrier(bar_id, GT);

    // ---- write the tile out, coalesced -------------------------------------------------------------------
    if (fused) {
      T* dst = a.score + (size_t)b_base * n_out;
      const int n_words = nq * n_out;
      if ((n_words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int i = gtid; i < n_words / 4; i += GT) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(os)[i];
      } else {
        for (int i = gtid; i < n_words; i += GT) dst[i] = os[i];
      }
    } else {
      if (a.write_score) {
        for (int i = gtid; i < nq * a.n_class; i += GT) {
          const int t = i / a.n_class, c = i - t * a.n_class;
          a.score[(size_t)(b_base + t) * a.score_ld + c] = os[t * n_out + c];
        }
      }
      if constexpr (MODE == M_GRAD) {
        T* gbase = a.grad + (a.jac_class > 0 ? (size_t)a.jac_class * a.n_in : 0);
        for (int i = gtid; i < nq * a.n_in; i += GT) {
          const int t = i / a.n_in, c = i - t * a.n_in;
          gbase[(size_t)(b_base + t) * a.grad_ld + c] = os[t * n_out + a.n_class + c];
        }
      }
    }
    group_barrier(bar_id, GT);  // xs / gs / qs / red are rewritten by the next tile
  }
}

// Host-side launcher for one instantiation.  Returns DC_ERR_UNSUPPORTED when the configuration does not fit shared
// memory (the caller falls back to the lane-split kernel).
template <int F, int KIND, int CW, int MODE, int NWG, int STAGES>
int launch_score_tq(ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  using Cfg = TqCfg<F, CW, MODE, NWG, STAGES>;
  constexpr size_t kMaxSmem = 227 * 1024;
  int ch = 32;
  while (Cfg::smem_bytes(ch) > kMaxSmem && ch > 4) ch >>= 1;  // shrink the ring chunk until the CTA fits
  if (Cfg::smem_bytes(ch) > kMaxSmem) return DC_ERR_UNSUPPORTED;
  a.st_q = Cfg::QT;
  a.chunk_rows = ch;
  a.n_tiles = (int)ceil_div64(a.batch, Cfg::QT);
  const size_t smem = Cfg::smem_bytes(ch);
  auto kern = score_tq_kernel<F, KIND, CW, MODE, NWG, STAGES>;
  DC_SET_FUNC_ATTR_ONCE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);  // per instantiation, per device
  const int grid = (int)min((long long)num_sms, (long long)ceil_div64(a.n_tiles, Cfg::NGRP));
  kern<<<grid, Cfg::NW * 32, smem, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc
