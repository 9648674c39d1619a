// dc_traj_step — one iteration of the reference's penalty trajectory optimiser (Weighted.step, diffco/optim.py:706-752;
// same terms as adam_traj_optimize, optim.py:86-127) for W waypoints in ONE launch of ONE CTA:
//
//   control points cp = FK(p)                                   (robot.fkine, model.py)
//   path length  sum |cp[i+1]-cp[i]|^2                          optim.py:715-717
//   max move     sum clamp(|cp[i+1,m]-cp[i,m]|^2 - v^2, 0)      optim.py:718-719
//   joint limit  sum clamp(lo-p,0) + clamp(p-hi,0)              optim.py:720-722
//   collision    sum clamp(score + bias, 0)                     optim.py:711-713 (score and d score/dp come from dc_score_grad)
//   gradient of  dif*length + w_mm*max_move + w_jl*joint_limit + w_col*collision  w.r.t. p, analytically (J_FK^T products),
//   masked, then torch.optim.Adam's update and robot.wrap.
//
// The reference builds this from ~40 tensor ops + autograd per step; the loop is launch-latency bound there
// (SURVEY.md §3.4).  Thread w owns waypoint w; control points are exchanged through shared memory; the four sums are
// reduced in a fixed order.
#include "dc_common.cuh"
#include "dc_fk.cuh"

namespace dc {

template <typename T>
struct TrajArgs {
  dc_fk_desc fk;
  dc_traj_params prm;
  T* p;
  const T* score;
  const T* score_grad;
  const T* mask;
  T* exp_avg;
  T* exp_avg_sq;
  double* step;
  T* terms;
  int n_wp;
  int n_feat;
  // dense collision checking (Weighted.step with options['dense_check'], optim.py:709-711): the collision term is taken
  // over the points of utils.dense_path (written by traj_dense_path_kernel, scored by dc_score_grad) instead of the waypoints
  const int* dense_count;   // nullptr: collision at the waypoints
  const int* seg_offset;    // [W]: first dense point of segment i; seg_offset[W-1] = index of the last point (= last waypoint)
  const T* dense_score;     // [max_points]
  const T* dense_grad;      // [max_points][dof]
  // device-side exit test (optim.py:747-752): state[0] != 0 -> this launch does nothing; after the update state[1] += 1 and
  // state[0] = 1 once the constraint loss is <= exit_constraint.  nullptr: never skip.
  int* state;
  double exit_constraint;
};

// utils.dense_path (utils.py:87-102) on the device, one CTA: per segment ceil(|dq| / max_step) points
// q[i] + k max_step dq / |dq| (k = 0, 1, ...), then the last waypoint.  Rows from the point count up to max_points are
// filled with the last waypoint so that the (static-size) scoring launch that follows reads defined values.
template <typename T>
__global__ void __launch_bounds__(256, 1) traj_dense_path_kernel(const T* __restrict__ p, int W, int D, T max_step, int max_points,
                                                                 T* __restrict__ dense, int* __restrict__ seg_offset,
                                                                 int* __restrict__ count, const int* __restrict__ state) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* cnt = reinterpret_cast<int*>(smem_raw);  // [W]: points per segment, then exclusive prefix
  if (state != nullptr && state[0] != 0) return;
  const int tid = threadIdx.x;
  for (int i = tid; i < W - 1; i += blockDim.x) {
    T ss = 0;
    for (int d = 0; d < D; ++d) {
      const T dl = p[(size_t)(i + 1) * D + d] - p[(size_t)i * D + d];
      ss += dl * dl;
    }
    const T dist = sqrt(ss);
    cnt[i] = (int)ceil(dist / max_step);
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < W - 1; ++i) {
      const int c = cnt[i];
      cnt[i] = run;
      run += c;
    }
    cnt[W - 1] = run;  // index of the final point
    *count = (run + 1 <= max_points) ? run + 1 : -1;
  }
  __syncthreads();
  const int total = cnt[W - 1] + 1;
  for (int i = tid; i < W; i += blockDim.x) seg_offset[i] = cnt[i];
  if (total > max_points) return;  // overflow: reported through *count, the host falls back
  for (int i = tid; i < W - 1; i += blockDim.x) {
    const int m0 = cnt[i], n = cnt[i + 1] - m0;
    if (n == 0) continue;
    T ss = 0;
    for (int d = 0; d < D; ++d) {
      const T dl = p[(size_t)(i + 1) * D + d] - p[(size_t)i * D + d];
      ss += dl * dl;
    }
    const T dist = sqrt(ss);
    for (int k = 0; k < n; ++k)
      for (int d = 0; d < D; ++d) {
        const T dl = p[(size_t)(i + 1) * D + d] - p[(size_t)i * D + d];
        dense[(size_t)(m0 + k) * D + d] = p[(size_t)i * D + d] + ((T)k * dl) * max_step / dist;  // utils.py:98
      }
  }
  for (int m = total - 1 + tid; m < max_points; m += blockDim.x)
    for (int d = 0; d < D; ++d) dense[(size_t)m * D + d] = p[(size_t)(W - 1) * D + d];
}


template <typename T>
__device__ __forceinline__ T wrap2pi_t(T th) {  // (pi + th) % (2 pi) - pi with Python's sign convention (utils.py:51-52)
  const T pi = (T)3.14159265358979323846, two_pi = (T)(2.0 * 3.14159265358979323846);
  const T a = pi + th;
  T m = fmod(a, two_pi);
  if (m != (T)0 && m < (T)0) m += two_pi;
  return m - pi;
}

template <typename T>
__global__ void __launch_bounds__(256, 1) traj_step_kernel(const __grid_constant__ TrajArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cps = reinterpret_cast<T*>(smem_raw);                  // [W][F]
  T* red = cps + (size_t)a.n_wp * a.n_feat;                 // [5][8] per-warp partial sums
  T* bseg = red + 40;                                       // [W][D] dense checking: what segment i hands to waypoint i + 1
  const int tid = threadIdx.x, W = a.n_wp, F = a.n_feat, D = a.fk.dof;
  if (a.state != nullptr && a.state[0] != 0) return;        // the exit test has fired: replays are no-ops
  const bool dense = a.dense_count != nullptr;
  const int n_dense = dense ? *a.dense_count : 0;
  const T dense_w = dense ? (T)W / (T)max(n_dense, 1) : (T)0;  // .mean() * len(p), optim.py:710-711
  const int M = a.fk.type == DC_FK_NONE ? F : a.fk.n_points, dim = a.fk.type == DC_FK_NONE ? 1 : a.fk.point_dim;
  const dc_traj_params& P = a.prm;
  const double step_d = *a.step + 1.0;

  for (int w = tid; w < W; w += blockDim.x) {
    T q[DC_MAX_DOF];
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) q[i] = (i < D) ? a.p[(size_t)w * D + i] : (T)0;
    fk_forward<T>(a.fk, q, cps + (size_t)w * F, 1);
  }
  __syncthreads();

  if (dense && P.collision_weight != 0) {
    // hinge on every dense point; its gradient w.r.t. the point goes to the two waypoints of the segment:
    // pt = q_i + k ms u, u = dq / |dq|:  d pt / d q_i = I - alpha (I - u u^T),  d pt / d q_{i+1} = alpha (I - u u^T),  alpha = k ms / |dq|
    for (int w = tid; w < W; w += blockDim.x) {
      for (int d = 0; d < D; ++d) bseg[(size_t)w * D + d] = (T)0;
      if (w == W - 1) continue;
      const int m0 = a.seg_offset[w], n = a.seg_offset[w + 1] - m0;
      if (n == 0) continue;
      T u[DC_MAX_DOF], ss = 0;
      for (int d = 0; d < D; ++d) {
        u[d] = a.p[(size_t)(w + 1) * D + d] - a.p[(size_t)w * D + d];
        ss += u[d] * u[d];
      }
      const T dist = sqrt(ss);
      for (int d = 0; d < D; ++d) u[d] /= dist;
      for (int k = 1; k < n; ++k) {  // k = 0 is the waypoint itself (alpha = 0)
        const int m = m0 + k;
        if (!(a.dense_score[m] + (T)P.safety_bias > (T)0)) continue;
        const T alpha = (T)k * (T)P.max_speed / dist;
        T uh = 0;
        for (int d = 0; d < D; ++d) uh += u[d] * a.dense_grad[(size_t)m * D + d];
        for (int d = 0; d < D; ++d) bseg[(size_t)w * D + d] += alpha * (a.dense_grad[(size_t)m * D + d] - u[d] * uh);
      }
    }
    __syncthreads();
  }
  const T v2 = (T)(P.max_speed * P.max_speed);
  const T bc1 = (T)(1.0 - pow(P.beta1, step_d)), bc2_sqrt = (T)sqrt(1.0 - pow(P.beta2, step_d));
  const T step_size = (T)P.lr / bc1;
  T acc_len = 0, acc_col = 0, acc_jl = 0, acc_mm = 0, acc_g2 = 0;
  for (int w = tid; w < W; w += blockDim.x) {
    T q[DC_MAX_DOF], gq[DC_MAX_DOF], gcp[DC_MAX_FEATURES];
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) {
      q[i] = (i < D) ? a.p[(size_t)w * D + i] : (T)0;
      gq[i] = (T)0;
    }
    for (int f = 0; f < F; ++f) gcp[f] = (T)0;
    const T* c0 = cps + (size_t)w * F;
    for (int m = 0; m < M; ++m) {
      if (w > 0) {
        const T* cm = c0 - F;
        T ss = 0;
        for (int k = 0; k < dim; ++k) {
          const T d = c0[m * dim + k] - cm[m * dim + k];
          ss = fma(d, d, ss);
        }
        const T coef = (T)2 * ((T)P.dif_weight + ((P.max_move_weight != 0 && ss - v2 > (T)0) ? (T)P.max_move_weight : (T)0));
        for (int k = 0; k < dim; ++k) gcp[m * dim + k] += coef * (c0[m * dim + k] - cm[m * dim + k]);
      }
      if (w < W - 1) {
        const T* cn = c0 + F;
        T ss = 0;
        for (int k = 0; k < dim; ++k) {
          const T d = cn[m * dim + k] - c0[m * dim + k];
          ss = fma(d, d, ss);
        }
        acc_len += ss;
        if (P.max_move_weight != 0) acc_mm += max(ss - v2, (T)0);
        const T coef = (T)2 * ((T)P.dif_weight + ((P.max_move_weight != 0 && ss - v2 > (T)0) ? (T)P.max_move_weight : (T)0));
        for (int k = 0; k < dim; ++k) gcp[m * dim + k] -= coef * (cn[m * dim + k] - c0[m * dim + k]);
      }
    }
    fk_vjp<T>(a.fk, q, const_cast<T*>(c0), 1, gcp, 1, gq);
    T hinge_on = (T)0;
    T gcol[DC_MAX_DOF];
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) gcol[i] = (T)0;
    if (dense) {
      if (P.collision_weight != 0) {
        // points of segment w (and the final point for the last waypoint): sum of hinges, sum of their gradients
        const int m0 = a.seg_offset[w], m1 = (w < W - 1) ? a.seg_offset[w + 1] : m0 + 1;
        for (int m = m0; m < m1; ++m) {
          const T hinge = a.dense_score[m] + (T)P.safety_bias;
          if (hinge > (T)0) {
            acc_col += hinge;
            for (int i = 0; i < D; ++i) gcol[i] += a.dense_grad[(size_t)m * D + i];
          }
        }
        const T cwm = (T)P.collision_weight * dense_w;
        for (int i = 0; i < D; ++i) {
          gcol[i] -= bseg[(size_t)w * D + i];
          if (w > 0) gcol[i] += bseg[(size_t)(w - 1) * D + i];
          gcol[i] *= cwm;
        }
      }
    } else if (a.score != nullptr && P.collision_weight != 0) {
      const T hinge = a.score[w] + (T)P.safety_bias;
      acc_col += max(hinge, (T)0);
      hinge_on = hinge > (T)0 ? (T)P.collision_weight : (T)0;
    }
    for (int i = 0; i < D; ++i) {
      const size_t idx = (size_t)w * D + i;
      T g = gq[i] + gcol[i];
      if (hinge_on != (T)0) g += hinge_on * a.score_grad[idx];
      if (P.joint_limit_weight != 0) {
        const T lo = (T)P.limits[i][0], hi = (T)P.limits[i][1];
        acc_jl += max(lo - q[i], (T)0) + max(q[i] - hi, (T)0);
        g += (T)P.joint_limit_weight * ((q[i] > hi ? (T)1 : (T)0) - (q[i] < lo ? (T)1 : (T)0));
      }
      if (a.mask != nullptr) g *= a.mask[idx];
      acc_g2 = fma(g, g, acc_g2);
      // torch.optim.Adam (no weight decay / amsgrad): lerp, addcmul, addcdiv
      T m1 = a.exp_avg[idx], m2 = a.exp_avg_sq[idx];
      m1 = m1 + (g - m1) * (T)(1.0 - P.beta1);
      m2 = m2 * (T)P.beta2 + (T)(1.0 - P.beta2) * g * g;
      a.exp_avg[idx] = m1;
      a.exp_avg_sq[idx] = m2;
      const T denom = sqrt(m2) / bc2_sqrt + (T)P.eps;
      T pn = q[i] - step_size * (m1 / denom);
      if (P.wrap[i]) pn = wrap2pi_t(pn);
      a.p[idx] = pn;
    }
  }
  // fixed-order block reduction of the four sums
  T vals[5] = {acc_len, acc_col, acc_jl, acc_mm, acc_g2};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    T v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[k * 8 + (tid >> 5)] = v;
  }
  __syncthreads();
  if (tid == 0) {
    T s[5];
    for (int k = 0; k < 5; ++k) {
      s[k] = 0;
      for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) s[k] += red[k * 8 + wp];
    }
    if (dense) s[1] *= dense_w;
    a.terms[0] = s[0];
    a.terms[1] = s[1];
    a.terms[2] = s[2];
    a.terms[3] = s[3];
    a.terms[4] = (T)P.collision_weight * s[1] + (T)P.max_move_weight * s[3] + (T)P.joint_limit_weight * s[2];
    a.terms[5] = s[4];  // |masked gradient|^2 (adam_traj_optimize's stopping test, optim.py:121-122)
    *a.step = step_d;
    if (a.state != nullptr) {
      a.state[1] += 1;
      if ((double)a.terms[4] <= a.exit_constraint) a.state[0] = 1;
    }
  }
}

template <typename T>
static int launch_traj_step(const dc_fk_desc* fk, const dc_traj_params* prm, int64_t n_wp, void* p, const void* score,
                            const void* score_grad, const dc_traj_dense* dense, const void* mask, void* exp_avg, void* exp_avg_sq,
                            double* step, void* terms, int32_t* state, double exit_constraint, cudaStream_t stream) {
  TrajArgs<T> a;
  a.fk = *fk;
  a.prm = *prm;
  a.p = static_cast<T*>(p);
  a.score = static_cast<const T*>(score);
  a.score_grad = static_cast<const T*>(score_grad);
  a.mask = static_cast<const T*>(mask);
  a.exp_avg = static_cast<T*>(exp_avg);
  a.exp_avg_sq = static_cast<T*>(exp_avg_sq);
  a.step = step;
  a.terms = static_cast<T*>(terms);
  a.n_wp = (int)n_wp;
  a.n_feat = fk->type == DC_FK_NONE ? fk->dof : fk->n_points * fk->point_dim;
  a.dense_count = dense ? dense->count : nullptr;
  a.seg_offset = dense ? dense->seg_offset : nullptr;
  a.dense_score = dense ? static_cast<const T*>(dense->score) : nullptr;
  a.dense_grad = dense ? static_cast<const T*>(dense->score_grad) : nullptr;
  a.state = state;
  a.exit_constraint = exit_constraint;
  const size_t smem = sizeof(T) * ((size_t)n_wp * a.n_feat + 40 + (size_t)n_wp * fk->dof);
  if (smem > 200 * 1024) return DC_ERR_UNSUPPORTED;
  auto kern = traj_step_kernel<T>;
  DC_SET_FUNC_ATTR_ONCE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  kern<<<1, 256, smem, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc

extern "C" int dc_traj_dense_path(const void* p, int64_t n_waypoints, int32_t dof, int32_t dtype, double max_step,
                                  int32_t max_points, void* dense, int32_t* seg_offset, int32_t* count, const int32_t* state,
                                  dc_stream_t stream) {
  if (!p || !dense || !seg_offset || !count || n_waypoints < 2 || n_waypoints > 12000 || dof < 1 || dof > DC_MAX_DOF)
    return DC_ERR_INVALID_ARG;
  if (!(max_step > 0) || max_points < n_waypoints) return DC_ERR_INVALID_ARG;
  const size_t smem = sizeof(int) * (size_t)n_waypoints;
  if (dtype == DC_F32)
    dc::traj_dense_path_kernel<float><<<1, 256, smem, (cudaStream_t)stream>>>((const float*)p, (int)n_waypoints, dof, (float)max_step,
                                                                             max_points, (float*)dense, seg_offset, count, state);
  else if (dtype == DC_F64)
    dc::traj_dense_path_kernel<double><<<1, 256, smem, (cudaStream_t)stream>>>((const double*)p, (int)n_waypoints, dof, max_step,
                                                                              max_points, (double*)dense, seg_offset, count, state);
  else
    return DC_ERR_INVALID_ARG;
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_traj_step(const dc_fk_desc* fk, const dc_traj_params* prm, int64_t n_waypoints, int32_t dtype, void* p,
                            const void* score, const void* score_grad, const void* mask, void* exp_avg, void* exp_avg_sq,
                            double* step, void* terms, dc_stream_t stream) {
  return dc_traj_step_ex(fk, prm, n_waypoints, dtype, p, score, score_grad, nullptr, mask, exp_avg, exp_avg_sq, step, terms, nullptr,
                         -1.0, stream);
}

extern "C" int dc_traj_step_ex(const dc_fk_desc* fk, const dc_traj_params* prm, int64_t n_waypoints, int32_t dtype, void* p,
                               const void* score, const void* score_grad, const dc_traj_dense* dense, const void* mask,
                               void* exp_avg, void* exp_avg_sq, double* step, void* terms, int32_t* state, double exit_constraint,
                               dc_stream_t stream) {
  if (!fk || !prm || !p || !exp_avg || !exp_avg_sq || !step || !terms) return DC_ERR_INVALID_ARG;
  if (n_waypoints < 2 || n_waypoints > (1 << 20) || fk->dof < 1 || fk->dof > DC_MAX_DOF) return DC_ERR_INVALID_ARG;
  if ((score == nullptr) != (score_grad == nullptr)) return DC_ERR_INVALID_ARG;
  if (dense && (!dense->count || !dense->seg_offset || !dense->score || !dense->score_grad || score != nullptr))
    return DC_ERR_INVALID_ARG;
  const int F = fk->type == DC_FK_NONE ? fk->dof : fk->n_points * fk->point_dim;
  if (F < 1 || F > DC_MAX_FEATURES) return DC_ERR_INVALID_ARG;
  if (!(prm->lr > 0) || !(prm->beta1 >= 0 && prm->beta1 < 1) || !(prm->beta2 >= 0 && prm->beta2 < 1)) return DC_ERR_INVALID_ARG;
  if (dtype == DC_F32)
    return dc::launch_traj_step<float>(fk, prm, n_waypoints, p, score, score_grad, dense, mask, exp_avg, exp_avg_sq, step, terms,
                                       state, exit_constraint, (cudaStream_t)stream);
  if (dtype == DC_F64)
    return dc::launch_traj_step<double>(fk, prm, n_waypoints, p, score, score_grad, dense, mask, exp_avg, exp_avg_sq, step, terms,
                                        state, exit_constraint, (cudaStream_t)stream);
  return DC_ERR_INVALID_ARG;
}
