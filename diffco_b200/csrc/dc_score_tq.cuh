// Fused collision-score kernel, thread-per-query form (large batches, fp32).
//
//   score[b,c] = sum_n w[n,c] k(|FK(q_b) - s_n|^2)         grad[b,:] = J_FK(q_b)^T sum_n omega_n coef_n (x_b - s_n)
//
// replaces, in ONE launch and without ever materialising the B x N kernel matrix, what the reference does with
// fkine -> torch.cdist -> pow/reciprocal -> matmul and the autograd backward of all of them
// (diffco/kernel_perceptrons.py:309-319,362-370; diffco/kernel.py:17-79; diffco/model.py:40-48,225-241).
//
// Work decomposition (B200: 148 SMs, one persistent NW-warp CTA per SM):
//   * the batch is cut into tiles of QT = 64 queries; CTA i owns a contiguous, balanced range of tiles;
//   * tiles are staged in super-tiles of up to NT = 32*NW queries: phase A runs FK one-query-per-thread into
//     shared memory (xs[F][ST], conflict-free columns), phase C runs the J^T product the same way;
//   * phase B, per tile: lane l of EVERY warp holds queries (2l, 2l+1) of the tile as the two halves of packed
//     registers, and the NW warps split the support set into NW contiguous slices, so all warps stay busy on any
//     batch that fills the SMs and the tail quantises at 64 queries;
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

template <int F, int CW, int MODE, int NW, int STAGES>
struct TqCfg {
  static constexpr int FPAD = round_up(F, 2);
  static constexpr int ROW = round_up(FPAD + CW, 4);
  static constexpr int QT = 64;
  static constexpr int NT = 32 * NW;
  static constexpr int NG = (MODE == M_GRAD) ? 1 : 0;
  static constexpr int NRED = CW + NG * F;
  static constexpr int BAR_BYTES = round_up(NW * STAGES * 8, 128);
  __host__ __device__ static constexpr size_t smem_bytes(int st_q, int chunk_rows) {
    return (size_t)BAR_BYTES + sizeof(float) * ((size_t)NW * STAGES * chunk_rows * ROW + (size_t)NW * NRED * QT +
                                                (size_t)F * st_q + (size_t)NG * F * st_q);
  }
};

template <int F, int KIND, int CW, int MODE, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32, 1) score_tq_kernel(const __grid_constant__ ScoreArgs<float> a) {
  using T = float;
  using Cfg = TqCfg<F, CW, MODE, NW, STAGES>;
  constexpr int FPAD = Cfg::FPAD, ROW = Cfg::ROW, QT = Cfg::QT, NT = Cfg::NT, NG = Cfg::NG, NRED = Cfg::NRED;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ST = a.st_q, CH = a.chunk_rows;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * STAGES;
  T* ring_all = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES);
  T* ring = ring_all + (size_t)warp * STAGES * CH * ROW;
  T* red = ring_all + (size_t)NW * STAGES * CH * ROW;
  T* xs = red + (size_t)NW * NRED * QT;
  T* gs = xs + (size_t)F * ST;

  // ---- this CTA's tiles and this warp's slice of the support set ---------------------------------
  const long long t0 = (long long)blockIdx.x * a.n_tiles / gridDim.x;
  const long long t1 = (long long)(blockIdx.x + 1) * a.n_tiles / gridDim.x;
  const int n0 = (int)((long long)warp * a.n_sv / NW);
  const int n1 = (int)((long long)(warp + 1) * a.n_sv / NW);
  const int n_chunks = (n1 - n0 + CH - 1) / CH;
  const long long total_chunks = (long long)n_chunks * (t1 - t0);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  auto issue = [&](long long gi) {  // lane 0 only: fetch chunk (gi % n_chunks) of the slice into stage gi % STAGES
    const int ci = (int)(gi % n_chunks);
    const int stage = (int)(gi % STAGES);
    const int r0 = n0 + ci * CH;
    const int rows = min(CH, n1 - r0);
    const uint32_t bytes = (uint32_t)rows * ROW * sizeof(T);
    mbar_expect_tx(&bars[stage], bytes);
    tma_bulk_g2s(ring + (size_t)stage * CH * ROW, a.table + (size_t)r0 * ROW, bytes, &bars[stage]);
  };
  if (lane == 0) {
    for (long long gi = 0; gi < STAGES && gi < total_chunks; ++gi) issue(gi);
  }
  long long gi = 0;  // chunks consumed so far by this warp

  const int tiles_per_st = ST / QT;
  for (long long st0 = t0; st0 < t1; st0 += tiles_per_st) {
    const int nt = (int)min((long long)tiles_per_st, t1 - st0);
    const long long b_base = st0 * QT;
    const int nq = (int)min((long long)nt * QT, a.batch - b_base);

    // ---- phase A: FK, one query per thread, features into xs[f][t] ---------------------------------
    if (tid < nt * QT) {
      if (tid < nq) {
        const T* qp = a.q + (size_t)(b_base + tid) * a.n_in;
        if (a.fk.type == DC_FK_NONE) {
#pragma unroll
          for (int f = 0; f < F; ++f) xs[(size_t)f * ST + tid] = qp[f];
        } else {
          T qv[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qp[i] : (T)0;
          fk_forward<T>(a.fk, qv, xs + tid, ST);
        }
      } else {
#pragma unroll
        for (int f = 0; f < F; ++f) xs[(size_t)f * ST + tid] = (T)0;
      }
    }
    __syncthreads();

    // ---- phase B: tiles ---------------------------------------------------------------------------
    for (int tl = 0; tl < nt; ++tl) {
      const int toff = tl * QT;
      P2 x[F];
      P2 sc[CW];
      P2 g[NG > 0 ? F : 1];
      P2 go[CW];
#pragma unroll
      for (int f = 0; f < F; ++f) x[f] = P2(*reinterpret_cast<const float2*>(xs + (size_t)f * ST + toff + 2 * lane));
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
          const long long b = b_base + toff + 2 * lane;
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

      for (int ci = 0; ci < n_chunks; ++ci, ++gi) {
        const int stage = (int)(gi % STAGES);
        const uint32_t parity = (uint32_t)((gi / STAGES) & 1);
        mbar_wait(&bars[stage], parity);
        const T* buf = ring + (size_t)stage * CH * ROW;
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
        if (lane == 0 && gi + STAGES < total_chunks) {
          fence_proxy_async();
          issue(gi + STAGES);
        }
      }

      // ---- combine the NW partial sums of every query (fixed order) ----------------------------------
      {
        float2* rq = reinterpret_cast<float2*>(red + (size_t)warp * NRED * QT) + lane;
#pragma unroll
        for (int c = 0; c < CW; ++c) rq[(size_t)c * (QT / 2)] = sc[c].v;
#pragma unroll
        for (int f = 0; f < NG * F; ++f) rq[(size_t)(CW + f) * (QT / 2)] = g[f].v;
      }
      __syncthreads();
      for (int idx = tid; idx < NRED * QT; idx += NT) {
        const int kk = idx / QT, qi = idx - kk * QT;
        T s = (T)0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[((size_t)w * NRED + kk) * QT + qi];
        if (kk < CW) {
          const long long b = b_base + toff + qi;
          if (kk < a.n_class && b < a.batch && a.write_score) a.score[(size_t)b * a.score_ld + kk] = s * a.rc.score_scale;
        } else {
          gs[(size_t)(kk - CW) * ST + toff + qi] = s * a.rc.grad_scale;
        }
      }
      __syncthreads();
    }

    // ---- phase C: J_FK^T g_x, one query per thread ---------------------------------------------------
    if constexpr (MODE == M_GRAD) {
      if (tid < nq) {
        const long long b = b_base + tid;
        const T* qp = a.q + (size_t)b * a.n_in;
        T scale = (T)1;
        if (CW == 1 && a.grad_out != nullptr) scale = a.grad_out[b];
        T* out = a.grad + (size_t)b * a.grad_ld + (a.jac_class > 0 ? (size_t)a.jac_class * a.n_in : 0);
        if (a.fk.type == DC_FK_NONE) {
#pragma unroll
          for (int f = 0; f < F; ++f) out[f] = scale * gs[(size_t)f * ST + tid];
        } else {
          T qv[DC_MAX_DOF], gq[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) {
            qv[i] = (i < a.n_in) ? qp[i] : (T)0;
            gq[i] = (T)0;
          }
          fk_vjp<T>(a.fk, qv, xs + tid, ST, gs + tid, ST, gq);
          for (int i = 0; i < a.n_in; ++i) out[i] = scale * gq[i];
        }
      }
      __syncthreads();
    }
  }
}

// Host-side launcher for one instantiation.  Returns DC_ERR_UNSUPPORTED when the configuration does not fit shared
// memory (the caller falls back to the lane-split kernel).
template <int F, int KIND, int CW, int MODE, int NW, int STAGES>
int launch_score_tq(ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  using Cfg = TqCfg<F, CW, MODE, NW, STAGES>;
  constexpr size_t kMaxSmem = 227 * 1024;
  int st_q = Cfg::NT, ch = 32;
  // shrink the ring chunk, then the super-tile, until the CTA fits
  while (Cfg::smem_bytes(st_q, ch) > kMaxSmem && ch > 4) ch >>= 1;
  while (Cfg::smem_bytes(st_q, ch) > kMaxSmem && st_q > Cfg::QT) st_q >>= 1;
  if (Cfg::smem_bytes(st_q, ch) > kMaxSmem) return DC_ERR_UNSUPPORTED;
  a.st_q = st_q;
  a.chunk_rows = ch;
  a.n_tiles = (int)ceil_div64(a.batch, Cfg::QT);
  const size_t smem = Cfg::smem_bytes(st_q, ch);
  auto kern = score_tq_kernel<F, KIND, CW, MODE, NW, STAGES>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    DC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    attr_set = true;
  }
  const int grid = (int)min((long long)num_sms, (long long)a.n_tiles);
  kern<<<grid, Cfg::NT, smem, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc
