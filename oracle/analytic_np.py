"""Closed-form (no autograd) float64 numpy statement of score + d(score)/dq.  TEST INFRASTRUCTURE ONLY.

Second, independent oracle: where ``diffco_oracle.py`` restates the reference literally (cdist + autograd),
this file evaluates the analytic formulas of SURVEY.md §9 that the CUDA kernels implement:

    score_c = sum_n w[n,c] k(rho_n)          rho_n = |x - s_n|^2,  x = FK(q)
    g_x     = sum_n omega_n 2k'(rho_n) (x - s_n)       omega_n = sum_c go_c w[n,c]
    g_q     = J_FK(q)^T g_x

It is checked against the autograd oracle in tests/test_oracle_vs_golden.py; pure-numpy loops, small cases.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------- radial kernels: k(rho), 2k'(rho)


def radial(kind: str, a: float, n: int, rho: np.ndarray):
    """Return (k, coef) with coef = 2 dk/drho, i.e. dk/dx = coef * (x - s).  kernel.py:17-79."""
    rho = np.asarray(rho, dtype=np.float64)
    if kind == "rq":  # kernel.py:24-25
        u = 1.0 / (1.0 + a / n * rho)
        return u**n, -2.0 * a * u ** (n + 1)
    if kind == "multiquadric":  # kernel.py:54
        k = np.sqrt(rho / a**2 + 1.0)
        return k, 1.0 / (a**2 * k)
    if kind == "polyharmonic":  # kernel.py:59-79 ; cdist sub-gradient is 0 at r == 0
        r = np.sqrt(rho)
        pos = rho > 0
        safe = np.where(pos, rho, 1.0)
        if n % 2 == 1:
            k = r**n / a
            coef = np.where(pos, n * safe ** (n / 2.0 - 1.0) / a, 0.0)
        else:
            k = np.where(pos, safe ** (n / 2.0) * 0.5 * np.log(safe) / a, 0.0)
            coef = np.where(pos, safe ** (n / 2.0 - 1.0) * (n * 0.5 * np.log(safe) + 1.0) / a, 0.0)
        return k, coef
    raise ValueError(kind)


def score_grad_features(x, S, W, kind, a, n, grad_out=None):
    """x (B,F), S (N,F), W (N,C) -> score (B,C), g_x (B,F) for upstream grad_out (B,C) (ones by default)."""
    x, S, W = (np.asarray(t, dtype=np.float64) for t in (x, S, W))
    if W.ndim == 1:
        W = W[:, None]
    B, C = x.shape[0], W.shape[1]
    go = np.ones((B, C)) if grad_out is None else np.asarray(grad_out, dtype=np.float64).reshape(B, C)
    score = np.zeros((B, C))
    gx = np.zeros_like(x)
    for b in range(B):
        d = x[b][None, :] - S
        k, coef = radial(kind, a, n, (d * d).sum(1))
        score[b] = k @ W
        omega = W @ go[b]
        gx[b] = (omega * coef) @ d
    return score, gx


# ---------------------------------------------------------------- FK maps with J^T products


def planar_chain(q, L):
    """model.py:40-48.  Returns x (B,D,2) and a closure vjp(gx)->(B,D) using the suffix-sum form."""
    q = np.asarray(q, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    th = np.cumsum(q, axis=1)
    px = np.cumsum(L * np.cos(th), axis=1)
    py = np.cumsum(L * np.sin(th), axis=1)
    pts = np.stack([px, py], axis=2)

    def vjp(g):
        g = g.reshape(pts.shape)
        prevx = np.concatenate([np.zeros((len(q), 1)), px[:, :-1]], axis=1)
        prevy = np.concatenate([np.zeros((len(q), 1)), py[:, :-1]], axis=1)
        # g_q[i] = sum_{j>=i} -gx_j (y_j - y_{i-1}) + gy_j (x_j - x_{i-1})
        sfx = lambda v: np.cumsum(v[:, ::-1], axis=1)[:, ::-1]
        A = sfx(-g[..., 0] * py + g[..., 1] * px)
        Gx, Gy = sfx(g[..., 0]), sfx(g[..., 1])
        return A + Gx * prevy - Gy * prevx

    return pts, vjp


def se2_body(q, kp):
    """model.py:90-93.  kp (2,M)."""
    q = np.asarray(q, dtype=np.float64)
    kp = np.asarray(kp, dtype=np.float64)
    c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
    rx = c[:, None] * kp[0] - s[:, None] * kp[1]
    ry = s[:, None] * kp[0] + c[:, None] * kp[1]
    pts = np.stack([rx + q[:, :1], ry + q[:, 1:2]], axis=2)

    def vjp(g):
        g = g.reshape(pts.shape)
        return np.stack([g[..., 0].sum(1), g[..., 1].sum(1), (-g[..., 0] * ry + g[..., 1] * rx).sum(1)], axis=1)

    return pts, vjp


def _rot_zyx(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz, Ry, Rx


def se3_body(q, kp):
    """model.py:156-159, utils.py:15-38.  kp (3,M); q=(x,y,z,roll,pitch,yaw)."""
    q = np.asarray(q, dtype=np.float64)
    kp = np.asarray(kp, dtype=np.float64)
    B, M = len(q), kp.shape[1]
    pts = np.zeros((B, M, 3))
    axes = np.zeros((B, 3, 3))  # rows: roll, pitch, yaw rotation axes in world frame
    for b in range(B):
        Rz, Ry, Rx = _rot_zyx(*q[b, 3:])
        pts[b] = (Rz @ Ry @ Rx @ kp).T + q[b, :3]
        axes[b, 0] = Rz @ Ry @ np.array([1.0, 0, 0])
        axes[b, 1] = Rz @ np.array([0, 1.0, 0])
        axes[b, 2] = np.array([0, 0, 1.0])

    def vjp(g):
        g = g.reshape(pts.shape)
        out = np.zeros((B, 6))
        out[:, :3] = g.sum(1)
        for b in range(B):
            rel = pts[b] - q[b, :3]
            for k in range(3):
                out[b, 3 + k] = (np.cross(axes[b, k][None, :], rel) * g[b]).sum()
        return out

    return pts, vjp


def _dh(theta, a, d, sa, ca):
    ct, st = np.cos(theta), np.sin(theta)
    return np.array([[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st], [0, sa, ca, d], [0, 0, 0, 1.0]])


def dh_multi(q, arms, interleave):
    """model.py:225-241, 366-383, 430-453, 486-503.  ``arms``: list of dicts with keys a,d,s_alpha,c_alpha,theta0,
    mask,joint_index and optional base (4,4), offset (3,), tool_points (T,3).  Revolute-joint columns:
    dp/dq_i = z_i x (p - o_i) with (z_i,o_i) taken from the frame *before* joint i."""
    q = np.asarray(q, dtype=np.float64)
    B, D = q.shape
    all_pts, all_dep = [], []  # per arm: list of (B,3) points ; per point: (arm, last joint idx it depends on)
    frames = []
    for arm in arms:
        ji = list(arm["joint_index"])
        J = len(ji)
        P = []
        dep = []
        Z = np.zeros((B, J, 3))
        O = np.zeros((B, J, 3))
        for b in range(B):
            T = np.eye(4) if arm.get("base") is None else np.asarray(arm["base"], dtype=np.float64)
            pb = []
            for i in range(J):
                Z[b, i], O[b, i] = T[:3, 2], T[:3, 3]
                T = T @ _dh(q[b, ji[i]] + float(arm["theta0"][i]), float(arm["a"][i]), float(arm["d"][i]),
                            float(arm["s_alpha"][i]), float(arm["c_alpha"][i]))
                if arm["mask"][i]:
                    pb.append(T[:3, 3].copy())
            tp = arm.get("tool_points")
            if tp is not None:
                for t in np.asarray(tp, dtype=np.float64):
                    pb.append((T @ np.append(t, 1.0))[:3])
            P.append(pb)
        P = np.array(P)  # (B,M,3)
        # NOTE: the offset translates outputs only; (z,o) and p shift together for o, so keep o un-offset
        # and subtract the offset from p when forming p - o.
        off = np.zeros(3) if arm.get("offset") is None else np.asarray(arm["offset"], dtype=np.float64)
        dep = [i for i in range(J) if arm["mask"][i]] + ([J - 1] * (0 if arm.get("tool_points") is None else len(arm["tool_points"])))
        all_pts.append(P + off)
        all_dep.append(dep)
        frames.append((Z, O, off, ji))
    if len(arms) == 1:
        pts = all_pts[0]
        order = [(0, m) for m in range(pts.shape[1])]
    elif interleave:
        M = all_pts[0].shape[1]
        order = [(a, m) for m in range(M) for a in range(len(arms))]
        pts = np.stack([all_pts[a][:, m] for a, m in order], axis=1)
    else:
        order = [(a, m) for a in range(len(arms)) for m in range(all_pts[a].shape[1])]
        pts = np.concatenate(all_pts, axis=1)

    def vjp(g):
        g = g.reshape(pts.shape)
        out = np.zeros((B, D))
        for k, (a, m) in enumerate(order):
            Z, O, off, ji = frames[a]
            p = all_pts[a][:, m] - off
            for i in range(all_dep[a][m] + 1):
                out[:, ji[i]] += (np.cross(Z[:, i], p - O[:, i]) * g[:, k]).sum(1)
        return out

    return pts, vjp


def se2_base_planar_arm(q, base_kp, L):
    """cfg-4 composed robot: SE(2) base keypoints followed by a planar chain expressed in the base frame."""
    q = np.asarray(q, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    base, base_vjp = se2_body(q[:, :3], base_kp)
    arm_l, arm_vjp = planar_chain(q[:, 3:], L)
    c, s = np.cos(q[:, 2])[:, None], np.sin(q[:, 2])[:, None]
    ax = c * arm_l[..., 0] - s * arm_l[..., 1]
    ay = s * arm_l[..., 0] + c * arm_l[..., 1]
    arm_w = np.stack([ax + q[:, :1], ay + q[:, 1:2]], axis=2)
    pts = np.concatenate([base, arm_w], axis=1)
    Mb = base.shape[1]

    def vjp(g):
        g = g.reshape(pts.shape)
        gb, ga = g[:, :Mb], g[:, Mb:]
        out = np.zeros_like(q)
        out[:, :3] = base_vjp(gb)
        out[:, 0] += ga[..., 0].sum(1)
        out[:, 1] += ga[..., 1].sum(1)
        out[:, 2] += (-ga[..., 0] * ay + ga[..., 1] * ax).sum(1)
        # rotate the upstream gradient back into the base frame, then the planar-chain vjp applies
        gl = np.stack([c * ga[..., 0] + s * ga[..., 1], -s * ga[..., 0] + c * ga[..., 1]], axis=2)
        out[:, 3:] = arm_vjp(gl)
        return out

    return pts, vjp
