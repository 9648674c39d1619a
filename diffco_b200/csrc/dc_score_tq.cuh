// Fused collision-score kernel, thread-per-query form (large batches, fp32).
//
//   score[b,c] = sum_n w[n,c] k(|FK(q_b) - s_n|^2)         grad[b,:] = J_FK(q_b)^T sum_n omega_n coef_n (x_b - s_n)
//
// replaces, in ONE launch and without ever materialising the B x N kernel matrix, what the reference does with
// fkine -> torch.cdist -> pow/reciprocal -> matmul and the autograd backward of all of them
// (diffco/kernel_perceptrons.py:309-319,362-370; diffco/kernel.py:17-79; diffco/model.py:40-48,225-241).
//
// Work decomposition (B200: 148 SMs, one persistent 16-warp CTA per SM):
//   * the batch is cut into tiles of QT = 32*Q queries; CTA i owns a contiguous, balanced range of tiles;
//   * tiles are staged in super-tiles of up to NT queries: phase A runs FK one-query-per-thread into shared
//     memory (xs[F][ST], conflict-free columns), phase C runs the J^T product the same way;
//   * phase B, per tile: lane l of EVERY warp holds the same Q queries in registers, and the NW warps split the
//     support set into NW contiguous slices, so all 16 warps stay busy on any batch size that fills the SMs and
//     the tail quantises at 32*Q queries instead of 512;
//   * each warp streams its own slice of the packed support table HBM/L2 -> shared memory with 1-D bulk TMA
//     (cp.async.bulk + mbarrier complete_tx) through a private STAGES-deep ring, running STAGES chunks ahead and
//     prefetching across tile boundaries; rows are read back as warp-uniform LDS.128 broadcasts;
//   * the pair update is 2-wide packed FP32 (FADD2/FFMA2): per support vector and query FP subtracts, FP fused
//     squares, one MUFU, ~6 scalar ops, FP fused gradient accumulations;
//   * partials are combined across warps through shared memory in a fixed order (deterministic results).
#pragma once

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

namespace dc {

enum ScoreMode { M_SCORE = 0, M_GRAD = 1, M_JAC = 2 };

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
  int n_feat;   // F
  int n_class;  // C
  int n_in;     // columns of q (dof, or F when fk.type == NONE)
  int n_tiles;  // ceil(batch / QT)
  int st_q;     // super-tile size in queries (multiple of QT, <= NT)
  int chunk_rows;
};

template <int FP, int CW, int MODE, int Q, int NW, int STAGES>
struct TqCfg {
  static constexpr int F2 = 2 * FP;
  static constexpr int ROW = round_up(F2 + CW, 4);
  static constexpr int QT = 32 * Q;
  static constexpr int NT = 32 * NW;
  static constexpr int NG = (MODE == M_JAC) ? CW : (MODE == M_GRAD ? 1 : 0);
  static constexpr int NRED = CW + NG * F2;
  static constexpr int BAR_BYTES = round_up(NW * STAGES * 8, 128);
  __host__ __device__ static constexpr size_t smem_bytes(int st_q, int chunk_rows) {
    return (size_t)BAR_BYTES + sizeof(float) * ((size_t)NW * STAGES * chunk_rows * ROW + (size_t)NW * NRED * QT +
                                                (size_t)F2 * st_q + (size_t)NG * F2 * st_q);
  }
};

template <int FP, int KIND, int CW, int MODE, int Q, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32, 1) score_tq_kernel(const __grid_constant__ ScoreArgs<float> a) {
  using T = float;
  using Cfg = TqCfg<FP, CW, MODE, Q, NW, STAGES>;
  constexpr int F2 = Cfg::F2, ROW = Cfg::ROW, QT = Cfg::QT, NT = Cfg::NT, NG = Cfg::NG, NRED = Cfg::NRED;
  using P = Pair<T>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ST = a.st_q, CH = a.chunk_rows;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * STAGES;
  T* ring_all = reinterpret_cast<T*>(smem_raw + Cfg::BAR_BYTES);
  T* ring = ring_all + (size_t)warp * STAGES * CH * ROW;
  T* red = ring_all + (size_t)NW * STAGES * CH * ROW;
  T* xs = red + (size_t)NW * NRED * QT;
  T* gs = xs + (size_t)F2 * ST;

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
          for (int f = 0; f < a.n_feat; ++f) xs[(size_t)f * ST + tid] = qp[f];
        } else {
          T qv[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qp[i] : (T)0;
          fk_forward<T>(a.fk, qv, xs + tid, ST);
        }
        for (int f = a.n_feat; f < F2; ++f) xs[(size_t)f * ST + tid] = (T)0;
      } else {
        for (int f = 0; f < F2; ++f) xs[(size_t)f * ST + tid] = (T)0;
      }
    }
    __syncthreads();

    // ---- phase B: tiles ---------------------------------------------------------------------------
    for (int tl = 0; tl < nt; ++tl) {
      const int toff = tl * QT;
      P x[Q][FP];
      T sc[Q][CW];
      P g[Q][NG > 0 ? NG : 1][FP];
      T go[Q][CW];
#pragma unroll
      for (int j = 0; j < Q; ++j) {
#pragma unroll
        for (int p = 0; p < FP; ++p)
          x[j][p] = P(xs[(size_t)(2 * p) * ST + toff + lane + 32 * j], xs[(size_t)(2 * p + 1) * ST + toff + lane + 32 * j]);
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          sc[j][c] = (T)0;
          go[j][c] = (T)1;
        }
#pragma unroll
        for (int gi2 = 0; gi2 < (NG > 0 ? NG : 1); ++gi2)
#pragma unroll
          for (int p = 0; p < FP; ++p) g[j][gi2][p] = P((T)0, (T)0);
        if constexpr (MODE == M_GRAD && CW > 1) {
          const long long b = b_base + toff + lane + 32 * j;
          if (a.grad_out != nullptr && b < a.batch) {
#pragma unroll
            for (int c = 0; c < CW; ++c) go[j][c] = (c < a.n_class) ? a.grad_out[(size_t)b * a.n_class + c] : (T)0;
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
#pragma unroll
          for (int j = 0; j < Q; ++j) {
            P d[FP];
            P acc0((T)0, (T)0), acc1((T)0, (T)0);
#pragma unroll
            for (int p = 0; p < FP; ++p) d[p] = padd(x[j][p], P(rowv[2 * p], rowv[2 * p + 1]));  // table holds -s
#pragma unroll
            for (int p = 0; p < FP; ++p) {
              if (p & 1)
                acc1 = pfma(d[p], d[p], acc1);
              else
                acc0 = pfma(d[p], d[p], acc0);
            }
            if constexpr (FP > 1) acc0 = padd(acc0, acc1);
            const T rho = acc0.lo() + acc0.hi();
            T k, coef;
            radial_eval<KIND, T>(a.rc, rho, k, coef);
            if constexpr (CW == 1) {
              const T w = rowv[F2];
              sc[j][0] = fma(w, k, sc[j][0]);
              if constexpr (MODE != M_SCORE) {
                const T c = w * coef;
                const P cc(c, c);
#pragma unroll
                for (int p = 0; p < FP; ++p) g[j][0][p] = pfma(cc, d[p], g[j][0][p]);
              }
            } else {
              T om = (T)0;
#pragma unroll
              for (int c = 0; c < CW; ++c) {
                const T w = rowv[F2 + c];
                sc[j][c] = fma(w, k, sc[j][c]);
                if constexpr (MODE == M_GRAD) om = fma(go[j][c], w, om);
                if constexpr (MODE == M_JAC) {
                  const T cv = w * coef;
                  const P cc(cv, cv);
#pragma unroll
                  for (int p = 0; p < FP; ++p) g[j][c][p] = pfma(cc, d[p], g[j][c][p]);
                }
              }
              if constexpr (MODE == M_GRAD) {
                const T cv = om * coef;
                const P cc(cv, cv);
#pragma unroll
                for (int p = 0; p < FP; ++p) g[j][0][p] = pfma(cc, d[p], g[j][0][p]);
              }
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
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        T* rq = red + (size_t)warp * NRED * QT + lane + 32 * j;
#pragma unroll
        for (int c = 0; c < CW; ++c) rq[(size_t)c * QT] = sc[j][c];
#pragma unroll
        for (int gi2 = 0; gi2 < NG; ++gi2)
#pragma unroll
          for (int p = 0; p < FP; ++p) {
            rq[(size_t)(CW + gi2 * F2 + 2 * p) * QT] = g[j][gi2][p].lo();
            rq[(size_t)(CW + gi2 * F2 + 2 * p + 1) * QT] = g[j][gi2][p].hi();
          }
      }
      __syncthreads();
      for (int idx = tid; idx < NRED * QT; idx += NT) {
        const int kk = idx / QT, qi = idx - kk * QT;
        T s = (T)0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[((size_t)w * NRED + kk) * QT + qi];
        if (kk < CW) {
          const long long b = b_base + toff + qi;
          if (kk < a.n_class && b < a.batch) a.score[(size_t)b * a.score_ld + kk] = s * a.rc.score_scale;
        } else {
          gs[(size_t)(kk - CW) * ST + toff + qi] = s * a.rc.grad_scale;
        }
      }
      __syncthreads();
    }

    // ---- phase C: J_FK^T g_x, one query per thread ---------------------------------------------------
    if constexpr (MODE != M_SCORE) {
      if (tid < nq) {
        const long long b = b_base + tid;
        const T* qp = a.q + (size_t)b * a.n_in;
        if (a.fk.type == DC_FK_NONE) {
          for (int gi2 = 0; gi2 < NG; ++gi2) {
            if (MODE == M_JAC && gi2 >= a.n_class) break;
            T scale = (T)1;
            if (MODE == M_GRAD && CW == 1 && a.grad_out != nullptr) scale = a.grad_out[b];
            T* out = a.grad + (size_t)b * a.grad_ld + (MODE == M_JAC ? (size_t)gi2 * a.n_in : 0);
            for (int f = 0; f < a.n_feat; ++f) out[f] = scale * gs[((size_t)gi2 * F2 + f) * ST + tid];
          }
        } else {
          T qv[DC_MAX_DOF];
#pragma unroll
          for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qp[i] : (T)0;
          for (int gi2 = 0; gi2 < NG; ++gi2) {
            if (MODE == M_JAC && gi2 >= a.n_class) break;
            T gq[DC_MAX_DOF];
#pragma unroll
            for (int i = 0; i < DC_MAX_DOF; ++i) gq[i] = (T)0;
            fk_vjp<T>(a.fk, qv, xs + tid, ST, gs + (size_t)gi2 * F2 * ST + tid, ST, gq);
            T scale = (T)1;
            if (MODE == M_GRAD && CW == 1 && a.grad_out != nullptr) scale = a.grad_out[b];
            T* out = a.grad + (size_t)b * a.grad_ld + (MODE == M_JAC ? (size_t)gi2 * a.n_in : 0);
            for (int i = 0; i < a.n_in; ++i) out[i] = scale * gq[i];
          }
        }
      }
      __syncthreads();
    }
  }
}

// Host-side launcher for one instantiation.  Returns false when the configuration does not fit shared memory.
template <int FP, int KIND, int CW, int MODE, int Q, int NW, int STAGES>
int launch_score_tq(ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  using Cfg = TqCfg<FP, CW, MODE, Q, NW, STAGES>;
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
  auto kern = score_tq_kernel<FP, KIND, CW, MODE, Q, NW, STAGES>;
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
