// Forward-kinematics feature maps and their Jacobian-transpose products, one query per thread.
//
// x = FK(q) follows diffco/model.py (== diffco/robot_fkine.py) and diffco/utils.py of the reference; the
// J^T products are the closed forms of SURVEY.md §9 (the reference gets them from autograd).  These run in
// the prologue / epilogue of the fused score kernel (once per query, against N support vectors in the main
// loop), so they are written for clarity and kept out of line; features live in a strided array (a shared
// memory column in the fused kernel, a global row in dc_fk_forward).
#pragma once

#include "dc_common.cuh"

namespace dc {

template <typename T>
struct Strided {
  T* p;
  int ld;
  __device__ __forceinline__ T& operator[](int i) const { return p[(size_t)i * ld]; }
};

// ---- planar revolute chain: model.py:40-48 -------------------------------------------------------
template <typename T>
__device__ void fk_planar_fwd(const double* link, int n, const T* q, Strided<T> x, T tx, T ty, T th0) {
  // The chain starts at (tx,ty) with heading th0 (zero for the plain chain; the SE(2) base pose for
  // DC_FK_SE2_BASE_PLANAR_ARM).  theta accumulates like torch.cumsum.
  T th = th0, px = tx, py = ty;
  for (int i = 0; i < n; ++i) {
    th += q[i];
    T s, c;
    sincos_t(th, &s, &c);
    px += (T)link[i] * c;
    py += (T)link[i] * s;
    x[2 * i] = px;
    x[2 * i + 1] = py;
  }
}

// The same chain with everything in registers (fully unrolled, at most NMAX links; x[2 NMAX], unused entries 0).  The
// expressions are those of fk_planar_fwd, so both produce bit-identical features.
template <int NMAX, typename T>
__device__ __forceinline__ void fk_planar_fwd_reg(const double* link, int n, const T* q, T* x) {
  T th = 0, px = 0, py = 0;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      th += q[i];
      T s, c;
      sincos_t(th, &s, &c);
      px += (T)link[i] * c;
      py += (T)link[i] * s;
      x[2 * i] = px;
      x[2 * i + 1] = py;
    } else {
      x[2 * i] = (T)0;
      x[2 * i + 1] = (T)0;
    }
  }
}

// g_q[i] = sum_{j>=i} -gx_j (y_j - y_{i-1}) + gy_j (x_j - x_{i-1}); p_{-1} = (ox, oy).
template <typename T>
__device__ void fk_planar_vjp(int n, Strided<T> x, Strided<T> g, T ox, T oy, T* gq) {
  T A = 0, Gx = 0, Gy = 0;
  for (int i = n - 1; i >= 0; --i) {
    T px = x[2 * i], py = x[2 * i + 1];
    T gx = g[2 * i], gy = g[2 * i + 1];
    A += gy * px - gx * py;
    Gx += gx;
    Gy += gy;
    T qx = (i > 0) ? x[2 * i - 2] : ox;
    T qy = (i > 0) ? x[2 * i - 1] : oy;
    gq[i] = A + Gx * qy - Gy * qx;
  }
}

// fk_planar_vjp with everything in registers (origin 0; fully unrolled, at most NMAX links; same expressions).
template <int NMAX>
__device__ __forceinline__ void fk_planar_vjp_reg(int n, const float* x, const float* g, float* gq) {
  float A = 0.f, Gx = 0.f, Gy = 0.f;
#pragma unroll
  for (int i = NMAX - 1; i >= 0; --i) {
    if (i < n) {
      const float px = x[2 * i], py = x[2 * i + 1];
      const float gx = g[2 * i], gy = g[2 * i + 1];
      A += gy * px - gx * py;
      Gx += gx;
      Gy += gy;
      const float qx = (i > 0) ? x[2 * (i > 0 ? i : 1) - 2] : 0.f;
      const float qy = (i > 0) ? x[2 * (i > 0 ? i : 1) - 1] : 0.f;
      gq[i] = A + Gx * qy - Gy * qx;
    }
  }
}

// ---- SE(2) rigid body: model.py:90-93, utils.py:40-48 ---------------------------------------------
template <typename T>
__device__ void fk_se2_fwd(const dc_fk_desc& fk, const T* q, Strided<T> x) {
  T s, c;
  sincos_t(q[2], &s, &c);
  for (int j = 0; j < fk.n_keypoints; ++j) {
    T kx = (T)fk.keypoints[0][j], ky = (T)fk.keypoints[1][j];
    x[2 * j] = c * kx - s * ky + q[0];
    x[2 * j + 1] = s * kx + c * ky + q[1];
  }
}
template <typename T>
__device__ void fk_se2_vjp(int m, const T* q, Strided<T> x, Strided<T> g, T* gq) {
  T g0 = 0, g1 = 0, g2 = 0;
  for (int j = 0; j < m; ++j) {
    T gx = g[2 * j], gy = g[2 * j + 1];
    g0 += gx;
    g1 += gy;
    g2 += gy * (x[2 * j] - q[0]) - gx * (x[2 * j + 1] - q[1]);
  }
  gq[0] = g0;
  gq[1] = g1;
  gq[2] = g2;
}

// ---- SE(3) rigid body: model.py:156-159, utils.py:15-38 (R = Rz(yaw) Ry(pitch) Rx(roll)) -----------
template <typename T>
__device__ void fk_se3_fwd(const dc_fk_desc& fk, const T* q, Strided<T> x) {
  T sr, cr, sp, cp, sy, cy;
  sincos_t(q[3], &sr, &cr);
  sincos_t(q[4], &sp, &cp);
  sincos_t(q[5], &sy, &cy);
  // Rz*Ry*Rx
  T r00 = cy * cp, r01 = cy * sp * sr - sy * cr, r02 = cy * sp * cr + sy * sr;
  T r10 = sy * cp, r11 = sy * sp * sr + cy * cr, r12 = sy * sp * cr - cy * sr;
  T r20 = -sp, r21 = cp * sr, r22 = cp * cr;
  for (int j = 0; j < fk.n_keypoints; ++j) {
    T kx = (T)fk.keypoints[0][j], ky = (T)fk.keypoints[1][j], kz = (T)fk.keypoints[2][j];
    x[3 * j] = r00 * kx + r01 * ky + r02 * kz + q[0];
    x[3 * j + 1] = r10 * kx + r11 * ky + r12 * kz + q[1];
    x[3 * j + 2] = r20 * kx + r21 * ky + r22 * kz + q[2];
  }
}
template <typename T>
__device__ void fk_se3_vjp(int m, const T* q, Strided<T> x, Strided<T> g, T* gq) {
  T f0 = 0, f1 = 0, f2 = 0, m0 = 0, m1 = 0, m2 = 0;  // force and moment (about the body origin) of g
  for (int j = 0; j < m; ++j) {
    T gx = g[3 * j], gy = g[3 * j + 1], gz = g[3 * j + 2];
    T rx = x[3 * j] - q[0], ry = x[3 * j + 1] - q[1], rz = x[3 * j + 2] - q[2];
    f0 += gx;
    f1 += gy;
    f2 += gz;
    m0 += ry * gz - rz * gy;
    m1 += rz * gx - rx * gz;
    m2 += rx * gy - ry * gx;
  }
  T sp, cp, sy, cy;
  sincos_t(q[4], &sp, &cp);
  sincos_t(q[5], &sy, &cy);
  gq[0] = f0;
  gq[1] = f1;
  gq[2] = f2;
  gq[3] = cy * cp * m0 + sy * cp * m1 - sp * m2;  // roll axis  Rz Ry e_x
  gq[4] = -sy * m0 + cy * m1;                     // pitch axis Rz e_y
  gq[5] = m2;                                     // yaw axis   e_z
}

// ---- serial standard-DH arms: utils.py:66-77, model.py:225-241 / 366-383 / 430-453 / 486-503 --------
// Affine frame [R|t] kept as 12 scalars.  `zo` (optional) receives, per joint, the axis z_i and origin o_i of
// the frame *before* joint i — what the revolute-joint Jacobian column z_i x (p - o_i) needs.
template <typename T>
__device__ void fk_dh_arm_fwd(const dc_dh_arm& arm, const T* q, Strided<T> x, T* zo) {
  T R[9], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) R[3 * r + c] = (T)arm.base[4 * r + c];
    t[r] = (T)arm.base[4 * r + 3];
  }
  const T ox = (T)arm.offset[0], oy = (T)arm.offset[1], oz = (T)arm.offset[2];
  for (int i = 0; i < arm.n_joints; ++i) {
    if (zo) {
      zo[6 * i + 0] = R[2];
      zo[6 * i + 1] = R[5];
      zo[6 * i + 2] = R[8];
      zo[6 * i + 3] = t[0];
      zo[6 * i + 4] = t[1];
      zo[6 * i + 5] = t[2];
    }
    T st, ct;
    sincos_t(q[arm.joint_index[i]] + (T)arm.theta0[i], &st, &ct);
    const T sa = (T)arm.s_alpha[i], ca = (T)arm.c_alpha[i], a = (T)arm.a[i], d = (T)arm.d[i];
    // DH_i = [[ct, -st ca, st sa, a ct], [st, ct ca, -ct sa, a st], [0, sa, ca, d]]
    const T m00 = ct, m01 = -st * ca, m02 = st * sa, m03 = a * ct;
    const T m10 = st, m11 = ct * ca, m12 = -ct * sa, m13 = a * st;
    const T m21 = sa, m22 = ca, m23 = d;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const T r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
      R[3 * r] = r0 * m00 + r1 * m10;
      R[3 * r + 1] = r0 * m01 + r1 * m11 + r2 * m21;
      R[3 * r + 2] = r0 * m02 + r1 * m12 + r2 * m22;
      t[r] = r0 * m03 + r1 * m13 + r2 * m23 + t[r];
    }
    const int slot = arm.out_slot[i];
    if (slot >= 0) {
      x[3 * slot] = t[0] + ox;
      x[3 * slot + 1] = t[1] + oy;
      x[3 * slot + 2] = t[2] + oz;
    }
  }
  for (int k = 0; k < arm.n_tool; ++k) {
    const T ux = (T)arm.tool[k][0], uy = (T)arm.tool[k][1], uz = (T)arm.tool[k][2];
    const int slot = arm.tool_slot[k];
    x[3 * slot] = R[0] * ux + R[1] * uy + R[2] * uz + t[0] + ox;
    x[3 * slot + 1] = R[3] * ux + R[4] * uy + R[5] * uz + t[1] + oy;
    x[3 * slot + 2] = R[6] * ux + R[7] * uy + R[8] * uz + t[2] + oz;
  }
}

// g_q[j_i] = z_i . ( sum_{p attached at frame >= i} (p - o_i) x g_p ) = z_i . (Mom - o_i x Frc)
template <typename T>
__device__ void fk_dh_arm_vjp(const dc_dh_arm& arm, const T* q, Strided<T> x, Strided<T> g, T* gq) {
  T zo[6 * DC_MAX_ARM_JOINTS];
  // Recompute the frames (cheaper than keeping 48 values alive across the pair loop).  This rewrites x with
  // the values it already holds (same thread, same inputs), which keeps a single forward code path.
  fk_dh_arm_fwd<T>(arm, q, x, zo);
  const T ox = (T)arm.offset[0], oy = (T)arm.offset[1], oz = (T)arm.offset[2];
  T F0 = 0, F1 = 0, F2 = 0, M0 = 0, M1 = 0, M2 = 0;
  auto add_point = [&](int slot) {
    const T px = x[3 * slot] - ox, py = x[3 * slot + 1] - oy, pz = x[3 * slot + 2] - oz;
    const T gx = g[3 * slot], gy = g[3 * slot + 1], gz = g[3 * slot + 2];
    F0 += gx;
    F1 += gy;
    F2 += gz;
    M0 += py * gz - pz * gy;
    M1 += pz * gx - px * gz;
    M2 += px * gy - py * gx;
  };
  for (int k = 0; k < arm.n_tool; ++k) add_point(arm.tool_slot[k]);
  for (int i = arm.n_joints - 1; i >= 0; --i) {
    if (arm.out_slot[i] >= 0) add_point(arm.out_slot[i]);
    const T zx = zo[6 * i], zy = zo[6 * i + 1], zz = zo[6 * i + 2];
    const T px = zo[6 * i + 3], py = zo[6 * i + 4], pz = zo[6 * i + 5];
    const T c0 = M0 - (py * F2 - pz * F1);
    const T c1 = M1 - (pz * F0 - px * F2);
    const T c2 = M2 - (px * F1 - py * F0);
    gq[arm.joint_index[i]] = zx * c0 + zy * c1 + zz * c2;
  }
}

// ---- URDF joint tree (collision_interfaces/rigid_body.py:86-141, urdf_interface.py:517-553) ------------------------
// Frames of all bodies, parents first: fr[i] = row-major [R | t] (12 values).
template <typename T>
__device__ void fk_tree_frames(const dc_fk_desc& fk, const T* q, T (*fr)[12]) {
  for (int i = 0; i < fk.n_nodes; ++i) {
    const dc_tree_node& nd = fk.tree[i];
    T Rp[9] = {(T)1, (T)0, (T)0, (T)0, (T)1, (T)0, (T)0, (T)0, (T)1}, tp[3] = {(T)0, (T)0, (T)0};
    if (nd.parent >= 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) Rp[e] = fr[nd.parent][e];
#pragma unroll
      for (int e = 0; e < 3; ++e) tp[e] = fr[nd.parent][9 + e];
    }
    const T qv = nd.q_index >= 0 ? (T)nd.mimic_mul * q[nd.q_index] + (T)nd.mimic_off : (T)0;
    // A = R_parent * rot
    T A[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        A[3 * r + c] = Rp[3 * r] * (T)nd.rot[c] + Rp[3 * r + 1] * (T)nd.rot[3 + c] + Rp[3 * r + 2] * (T)nd.rot[6 + c];
    T tj[3] = {(T)nd.trans[0], (T)nd.trans[1], (T)nd.trans[2]};
    T* o = fr[i];
    if (nd.joint >= DC_JOINT_REV_X && nd.joint <= DC_JOINT_REV_Z) {
      T sn, cs;
      sincos_t((T)nd.axis[0] * qv, &sn, &cs);
      // A * Rot_k(angle): the two columns other than k mix; (u, v) = the next two axes in cyclic order
      const int k = nd.joint - DC_JOINT_REV_X, u = (k + 1) % 3, v = (k + 2) % 3;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const T au = A[3 * r + u], av = A[3 * r + v];
        o[3 * r + k] = A[3 * r + k];
        o[3 * r + u] = au * cs + av * sn;
        o[3 * r + v] = av * cs - au * sn;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) o[e] = A[e];
      if (nd.joint == DC_JOINT_PRISMATIC) {
        // trans + rot * axis * q'
#pragma unroll
        for (int r = 0; r < 3; ++r)
          tj[r] += ((T)nd.rot[3 * r] * (T)nd.axis[0] + (T)nd.rot[3 * r + 1] * (T)nd.axis[1] + (T)nd.rot[3 * r + 2] * (T)nd.axis[2]) * qv;
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) o[9 + r] = tp[r] + Rp[3 * r] * tj[0] + Rp[3 * r + 1] * tj[1] + Rp[3 * r + 2] * tj[2];
  }
}

template <typename T>
__device__ __noinline__ void fk_tree_fwd(const dc_fk_desc& fk, const T* q, Strided<T> x) {
  T fr[DC_MAX_TREE_NODES][12];
  fk_tree_frames<T>(fk, q, fr);
  for (int i = 0; i < fk.n_nodes; ++i) {
    const int slot = fk.tree[i].out_slot;
    if (slot >= 0) {
      x[3 * slot] = fr[i][9];
      x[3 * slot + 1] = fr[i][10];
      x[3 * slot + 2] = fr[i][11];
    }
  }
}

// gq = J^T g.  With F_i / M_i the sums of g_l and p_l x g_l over the bodies l of subtree(i):
//   revolute:  d p_l / d q' = a_i x (p_l - t_i),  a_i = sign * column k of R_i   ->  gq' = a_i . (M_i - t_i x F_i)
//   prismatic: d p_l / d q' = R_i axis                                           ->  gq' = (R_i axis) . F_i
template <typename T>
__device__ __noinline__ void fk_tree_vjp(const dc_fk_desc& fk, const T* q, Strided<T> g, T* gq) {
  T fr[DC_MAX_TREE_NODES][12];
  T acc[DC_MAX_TREE_NODES][6];
  fk_tree_frames<T>(fk, q, fr);
  for (int i = 0; i < fk.n_nodes; ++i) {
#pragma unroll
    for (int e = 0; e < 6; ++e) acc[i][e] = (T)0;
    if (fk.tree[i].q_index >= 0) gq[fk.tree[i].q_index] = (T)0;
  }
  for (int i = fk.n_nodes - 1; i >= 0; --i) {
    const dc_tree_node& nd = fk.tree[i];
    const T* f = fr[i];
    if (nd.out_slot >= 0) {
      const T gx = g[3 * nd.out_slot], gy = g[3 * nd.out_slot + 1], gz = g[3 * nd.out_slot + 2];
      acc[i][0] += gx;
      acc[i][1] += gy;
      acc[i][2] += gz;
      acc[i][3] += f[10] * gz - f[11] * gy;
      acc[i][4] += f[11] * gx - f[9] * gz;
      acc[i][5] += f[9] * gy - f[10] * gx;
    }
    if (nd.q_index >= 0) {
      T d = (T)0;
      if (nd.joint == DC_JOINT_PRISMATIC) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
          d += (f[3 * r] * (T)nd.axis[0] + f[3 * r + 1] * (T)nd.axis[1] + f[3 * r + 2] * (T)nd.axis[2]) * acc[i][r];
      } else if (nd.joint != DC_JOINT_FIXED) {
        const int k = nd.joint - DC_JOINT_REV_X;
        const T ax = f[k], ay = f[3 + k], az = f[6 + k];
        const T mx = acc[i][3] - (f[10] * acc[i][2] - f[11] * acc[i][1]);
        const T my = acc[i][4] - (f[11] * acc[i][0] - f[9] * acc[i][2]);
        const T mz = acc[i][5] - (f[9] * acc[i][1] - f[10] * acc[i][0]);
        d = (T)nd.axis[0] * (ax * mx + ay * my + az * mz);
      }
      gq[nd.q_index] += (T)nd.mimic_mul * d;
    }
    if (nd.parent >= 0) {
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[nd.parent][e] += acc[i][e];
    }
  }
}

// ---- dispatch ----------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void fk_forward_one(const dc_fk_desc& fk, const T* q, T* xp, int ld) {
  Strided<T> x{xp, ld};
  switch (fk.type) {
    case DC_FK_NONE:
      for (int i = 0; i < fk.dof; ++i) x[i] = q[i];
      break;
    case DC_FK_PLANAR_CHAIN:
      fk_planar_fwd<T>(fk.link_length, fk.n_links, q, x, (T)0, (T)0, (T)0);
      break;
    case DC_FK_SE2_BODY:
      fk_se2_fwd<T>(fk, q, x);
      break;
    case DC_FK_SE3_BODY:
      fk_se3_fwd<T>(fk, q, x);
      break;
    case DC_FK_DH_ARMS:
      for (int a = 0; a < fk.n_arms; ++a) fk_dh_arm_fwd<T>(fk.arms[a], q, x, nullptr);
      break;
    case DC_FK_SE2_BASE_PLANAR_ARM: {
      fk_se2_fwd<T>(fk, q, x);
      // chain in the base frame == chain with initial angle theta and origin (x, y)
      Strided<T> xa{xp + (size_t)2 * fk.n_keypoints * ld, ld};
      fk_planar_fwd<T>(fk.link_length, fk.n_links, q + 3, xa, q[0], q[1], q[2]);
      break;
    }
    case DC_FK_JOINT_TREE:
      fk_tree_fwd<T>(fk, q, x);
      break;
    default:
      break;
  }
}

// Composite maps (dc_fk_desc.n_repeat / time_last): the map above applied to n_repeat consecutive blocks of q with the
// features concatenated (LineFKKernel, kernel.py:145-173) and / or a trailing time column passed through as the last
// feature (TemporalFKKernel, kernel.py:175-202).
__device__ __forceinline__ int fk_total_features(const dc_fk_desc& fk) {
  return fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
}
__device__ __forceinline__ bool fk_composite(const dc_fk_desc& fk) { return fk.n_repeat > 1 || fk.time_last != 0; }

template <typename T>
__device__ __noinline__ void fk_forward(const dc_fk_desc& fk, const T* q, T* xp, int ld) {
  const int rep = fk.n_repeat > 1 ? fk.n_repeat : 1, tl = fk.time_last ? 1 : 0;
  const int din = (fk.dof - tl) / rep, fin = (fk_total_features(fk) - tl) / rep;
  for (int r = 0; r < rep; ++r) fk_forward_one<T>(fk, q + r * din, xp + (size_t)r * fin * ld, ld);
  if (tl) xp[(size_t)rep * fin * ld] = q[fk.dof - 1];
}

// sin / cos in float64 to ~1e-13 absolute — all the (hi, lo) float32 feature pairs need (lo is ~1e-7 of hi) — at a
// fifth of the instruction count of libm's sincos: two-term Cody-Waite reduction to [-pi/4, pi/4], Taylor polynomials
// (sin to x^15, cos to x^14: truncation < 5e-13 at pi/4).  Arguments here are sums of joint angles (|x| < a few hundred).
// Polynomial coefficients live in constant memory: a DFMA takes them as c[bank][offset] operands, whereas literals are
// materialised as two 32-bit moves each — a fifth of the instructions of the unrolled planar chain.
static __constant__ double kSinCos64[19] = {
    0.63661977236758134308, 1.57079632673412561417, 6.07710050650619224932e-11,
    1.0 / 1307674368000.0, -1.0 / 6227020800.0, 1.0 / 39916800.0, -1.0 / 362880.0, 1.0 / 5040.0, -1.0 / 120.0, 1.0 / 6.0,
    -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5, 1.0, 0.0};
__device__ __forceinline__ void sincos_fast64(double x, double* s, double* c) {
  const double* K = kSinCos64;
  const double kd = rint(x * K[0]);          // x / (pi/2)
  double r = fma(-kd, K[1], x);              // pi/2 = P1 + P2, P1 exact in 33 bits
  r = fma(-kd, K[2], r);
  const double r2 = r * r;
  double sp = K[3];                          // 1/15!
  sp = fma(sp, r2, K[4]);
  sp = fma(sp, r2, K[5]);
  sp = fma(sp, r2, K[6]);
  sp = fma(sp, r2, K[7]);
  sp = fma(sp, r2, K[8]);
  sp = fma(sp, r2, K[9]);
  sp = fma(-sp * r2, r, r);                  // r - r^3 (1/6 - ...)
  double cp = K[10];                         // -1/14!
  cp = fma(cp, r2, K[11]);
  cp = fma(cp, r2, K[12]);
  cp = fma(cp, r2, K[13]);
  cp = fma(cp, r2, K[14]);
  cp = fma(cp, r2, K[15]);
  cp = fma(cp, r2, K[16]);
  cp = fma(cp, r2, K[17]);
  const int k = (int)kd;
  const double s0 = (k & 1) ? cp : sp, c0 = (k & 1) ? sp : cp;
  *s = (k & 2) ? -s0 : s0;
  *c = ((k + 1) & 2) ? -c0 : c0;
}

// The planar chain in float64 with sincos_fast64, everything in registers: THE evaluation behind the float32 features of
// RevolutePlanarRobot (tensor-core kernel's tile prologue, fk_forward_f32x) — one function, so features are bit-identical
// wherever they are computed.
template <int NMAX>
__device__ __forceinline__ void fk_planar_f64_reg(const double* link, int n, const float* q, double* x) {
  double th = 0.0, px = 0.0, py = 0.0;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      th += (double)q[i];
      double s, c;
      sincos_fast64(th, &s, &c);
      px = fma(link[i], c, px);
      py = fma(link[i], s, py);
      x[2 * i] = px;
      x[2 * i + 1] = py;
    } else {
      x[2 * i] = 0.0;
      x[2 * i + 1] = 0.0;
    }
  }
}

// float32 FEATURES from a float64 evaluation of the map on the same float32 configuration: hi = fl32(x), and (WITH_LO)
// lo = fl32(x - hi).  A float32 chain accumulates ~1e-6 of absolute error over seven joints (angle sums up to 7 pi,
// positions up to 7 link lengths); next to a support vector that alone is several 1e-5 of the gradient maximum for ANY
// float32 evaluation (tests/test_gpu_tc_stress.py).  Evaluating in float64 leaves only the final rounding (hi), and the
// exact near-pair path of the tensor-core kernel adds lo back, so differences x - s are good to ~1e-9.  Every float32
// kernel and dc_fk_forward use THIS function, so a query that coincides with a support has bit-identical features.
template <bool WITH_LO>
__device__ __noinline__ void fk_forward_f32x(const dc_fk_desc& fk, const float* q, float* xh, int ldh, float* xl, int ldl) {
  double qd[DC_MAX_DOF], xd[DC_MAX_FEATURES];
  if (fk.type == DC_FK_PLANAR_CHAIN && fk.n_links <= 8 && !fk_composite(fk)) {
    fk_planar_f64_reg<8>(fk.link_length, fk.n_links, q, xd);
  } else {
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) qd[i] = (i < fk.dof) ? (double)q[i] : 0.0;
    fk_forward<double>(fk, qd, xd, 1);
  }
  const int F = fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
  for (int f = 0; f < F; ++f) {
    const float h = (float)xd[f];
    xh[(size_t)f * ldh] = h;
    if (WITH_LO) xl[(size_t)f * ldl] = (float)(xd[f] - (double)h);
  }
}
// the feature map as the kernels of element type T evaluate it
__device__ __forceinline__ void fk_features(const dc_fk_desc& fk, const float* q, float* x, int ld) {
  if (fk.type == DC_FK_NONE) {
    for (int i = 0; i < fk.dof; ++i) x[(size_t)i * ld] = q[i];
  } else {
    fk_forward_f32x<false>(fk, q, x, ld, nullptr, 0);
  }
}
__device__ __forceinline__ void fk_features(const dc_fk_desc& fk, const double* q, double* x, int ld) {
  fk_forward<double>(fk, q, x, ld);
}

template <typename T>
__device__ __forceinline__ void fk_vjp_one(const dc_fk_desc& fk, const T* q, T* xp, int ldx, T* gp, int ldg, T* gq) {
  Strided<T> x{xp, ldx};
  Strided<T> g{gp, ldg};
  switch (fk.type) {
    case DC_FK_NONE:
      for (int i = 0; i < fk.dof; ++i) gq[i] = g[i];
      break;
    case DC_FK_PLANAR_CHAIN:
      fk_planar_vjp<T>(fk.n_links, x, g, (T)0, (T)0, gq);
      break;
    case DC_FK_SE2_BODY:
      fk_se2_vjp<T>(fk.n_keypoints, q, x, g, gq);
      break;
    case DC_FK_SE3_BODY:
      fk_se3_vjp<T>(fk.n_keypoints, q, x, g, gq);
      break;
    case DC_FK_DH_ARMS:
      for (int a = 0; a < fk.n_arms; ++a) fk_dh_arm_vjp<T>(fk.arms[a], q, x, g, gq);
      break;
    case DC_FK_SE2_BASE_PLANAR_ARM: {
      const int mb = fk.n_keypoints;
      fk_se2_vjp<T>(mb, q, x, g, gq);
      Strided<T> xa{xp + (size_t)2 * mb * ldx, ldx};
      Strided<T> ga{gp + (size_t)2 * mb * ldg, ldg};
      // The arm points rotate with the base: they contribute to (x, y, theta) exactly like body key points,
      // and to the arm joints through the planar-chain formula with p_{-1} = base origin.
      T a3[3];
      fk_se2_vjp<T>(fk.n_links, q, xa, ga, a3);
      gq[0] += a3[0];
      gq[1] += a3[1];
      gq[2] += a3[2];
      fk_planar_vjp<T>(fk.n_links, xa, ga, q[0], q[1], gq + 3);
      break;
    }
    case DC_FK_JOINT_TREE:
      fk_tree_vjp<T>(fk, q, g, gq);
      break;
    default:
      break;
  }
}

template <typename T>
__device__ __noinline__ void fk_vjp(const dc_fk_desc& fk, const T* q, T* xp, int ldx, T* gp, int ldg, T* gq) {
  const int rep = fk.n_repeat > 1 ? fk.n_repeat : 1, tl = fk.time_last ? 1 : 0;
  const int din = (fk.dof - tl) / rep, fin = (fk_total_features(fk) - tl) / rep;
  for (int r = 0; r < rep; ++r)
    fk_vjp_one<T>(fk, q + r * din, xp + (size_t)r * fin * ldx, ldx, gp + (size_t)r * fin * ldg, ldg, gq + r * din);
  if (tl) gq[fk.dof - 1] = gp[(size_t)rep * fin * ldg];
}

}  // namespace dc
