"""CPU oracle for the DiffCo collision-score hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

A functional restatement (plain PyTorch CPU ops, float64 by default) of the reference algorithm for the
path named in BASELINE.json: FK feature map -> radial kernel vs. support vectors -> weighted sum, with the
gradient obtained exactly the way the reference obtains it (autograd through FK, ``torch.cdist`` and the
kernel formula).  Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; ``diffco_b200/`` never does.

Pinning: ``oracle/make_golden.py`` runs the *unmodified* reference (imported through
``oracle/ref_loader.py``) and this restatement on identical seeded inputs and commits the reference's
outputs under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` re-checks this file against those
fixtures on every run (and against the live reference when ``/root/reference`` exists).  The reference
itself ships no golden vectors for this path, so reference-generated fixtures are the
pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch

# --------------------------------------------------------------------------------------------------
# Radial kernels (diffco/kernel.py)
# --------------------------------------------------------------------------------------------------


def _flat_rows(xs: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    # kernel.py:18-21 — prepend dims until ranks match, then flatten everything after dim 0.
    while xs.ndim < like.ndim:
        xs = xs.unsqueeze(0)
    return xs.reshape(xs.shape[0], -1)


def rq_kernel(xs, x_primes, gamma: float, p: int = 2):
    """RQKernel.__call__, kernel.py:17-29: (1 + gamma/p * r^2)^-p with r from torch.cdist; rows==1 squeezed."""
    a = _flat_rows(xs, x_primes)
    b = x_primes.reshape(x_primes.shape[0], -1)
    r2 = torch.cdist(a, b).square()
    k = 1 / (1 + gamma / p * r2) ** p
    return k.squeeze(0) if k.shape[0] == 1 else k


def temporal_fk_kernel(xs, x_primes, fkine, gamma, p, gamma_t, p_t, alpha):
    """TemporalFKKernel.__call__, kernel.py:182-195: rows are [q | t]; rq(FK(q), FK(q')) * t_rq(t, t') ** alpha."""
    if xs.ndim == 1:
        xs = xs[None, :]
    xc = fkine(xs[:, :-1]).reshape(len(xs), -1)
    sc = fkine(x_primes[:, :-1]).reshape(len(x_primes), -1)
    return rq_kernel(xc, sc, gamma, p) * rq_kernel(xs[:, -1:], x_primes[:, -1:], gamma_t, p_t) ** alpha


def line_fk_kernel(xs, x_primes, fkine, gamma, p: int = 2):
    """LineFKKernel.__call__, kernel.py:209-220: rows are [q_a | q_b]; rq on the concatenated features of both ends."""
    if xs.ndim == 1:
        xs = xs[None, :]
    if x_primes.ndim == 1:
        x_primes = x_primes[None, :]
    dof = xs.shape[1] // 2
    xc = fkine(xs.reshape(-1, dof)).reshape(len(xs), -1)
    sc = fkine(x_primes.reshape(-1, dof)).reshape(len(x_primes), -1)
    return rq_kernel(xc, sc, gamma, p)


def line_kernel(xs, x_primes, point_kernel):
    """LineKernel.__call__, kernel.py:197-207: mean of the point kernel on the two halves of [q_a | q_b] rows."""
    if xs.ndim == 1:
        xs = xs[None, :]
    if x_primes.ndim == 1:
        x_primes = x_primes[None, :]
    dof = xs.shape[1] // 2
    return (point_kernel(xs[:, :dof], x_primes[:, :dof]) + point_kernel(xs[:, dof:], x_primes[:, dof:])) / 2


def polyharmonic_kernel(xs, x_primes, k: int, epsilon: float):
    """Polyharmonic.__call__, kernel.py:59-79: r^k/eps (k odd) or r^k log r/eps with NaN->0 (k even); never squeezed."""
    a = _flat_rows(xs, x_primes)
    b = x_primes.reshape(x_primes.shape[0], -1)
    r = torch.cdist(a, b)
    if k % 2 == 0:
        v = r**k * torch.log(r)
        v = torch.where(torch.isnan(v), torch.zeros_like(v), v)
    else:
        v = r if k == 1 else r**k
    return v / epsilon


def multiquadric_kernel(xs, x_primes, epsilon: float):
    """MultiQuadratic.__call__, kernel.py:45-57: sqrt(|x-s|^2/eps^2 + 1) by explicit broadcast difference."""
    if xs.ndim == 1:
        xs = xs[None, :]
    diff = x_primes[None, :, :] - xs[:, None, :]
    k = torch.sqrt((diff**2).sum(dim=2) / epsilon**2 + 1)
    return k.squeeze(0) if k.shape[0] == 1 else k


@dataclass
class KernelSpec:
    """kind in {'rq','polyharmonic','multiquadric'}; a = gamma | epsilon ; n = p | k."""

    kind: str
    a: float
    n: int = 2

    def __call__(self, xs, x_primes):
        if self.kind == "rq":
            return rq_kernel(xs, x_primes, self.a, self.n)
        if self.kind == "polyharmonic":
            return polyharmonic_kernel(xs, x_primes, self.n, self.a)
        if self.kind == "multiquadric":
            return multiquadric_kernel(xs.reshape(xs.shape[0], -1) if xs.ndim > 1 else xs,
                                       x_primes.reshape(x_primes.shape[0], -1), self.a)
        raise ValueError(self.kind)


# --------------------------------------------------------------------------------------------------
# Forward kinematics feature maps (diffco/model.py == diffco/robot_fkine.py, diffco/utils.py)
# --------------------------------------------------------------------------------------------------


def fk_planar_chain(q, link_length):
    """RevolutePlanarRobot.fkine, model.py:40-48: theta=cumsum(q); (x,y)=cumsum(L*cos/sin theta) -> (B,D,2)."""
    L = torch.as_tensor(link_length)
    q = q.reshape(-1, L.numel())
    th = torch.cumsum(q, dim=1)
    return torch.stack([torch.cumsum(L * th.cos(), 1), torch.cumsum(L * th.sin(), 1)], dim=2)


def fk_se2_body(q, keypoints):
    """RigidPlanarBody.fkine, model.py:90-93 with utils.rot_2d, utils.py:40-48.  keypoints: (2,M) -> (B,M,2)."""
    kp = torch.as_tensor(keypoints)
    q = q.reshape(-1, 3)
    c, s = q[:, 2].cos(), q[:, 2].sin()
    R = torch.stack([torch.stack([c, -s], 1), torch.stack([s, c], 1)], 1)  # (B,2,2)
    pts = R.to(q.dtype) @ kp.to(q.dtype) + q[:, :2, None]
    return pts.permute(0, 2, 1)


def euler_zyx(phi):
    """utils.euler2mat, utils.py:15-38: Rz(yaw) @ Ry(pitch) @ Rx(roll) for phi=(roll,pitch,yaw)."""
    phi = phi.reshape(-1, 3)
    s, c = phi.sin(), phi.cos()
    one, zero = torch.ones_like(s[:, 0]), torch.zeros_like(s[:, 0])

    def m(rows):
        return torch.stack([torch.stack(r, 1) for r in rows], 1)

    rx = m([[one, zero, zero], [zero, c[:, 0], -s[:, 0]], [zero, s[:, 0], c[:, 0]]])
    ry = m([[c[:, 1], zero, s[:, 1]], [zero, one, zero], [-s[:, 1], zero, c[:, 1]]])
    rz = m([[c[:, 2], -s[:, 2], zero], [s[:, 2], c[:, 2], zero], [zero, zero, one]])
    return rz @ ry @ rx


def fk_se3_body(q, keypoints):
    """RigidBody.fkine, model.py:156-159.  keypoints: (3,M) -> (B,M,3); q=(x,y,z,roll,pitch,yaw)."""
    kp = torch.as_tensor(keypoints)
    q = q.reshape(-1, 6)
    pts = euler_zyx(q[:, 3:]) @ kp.to(q.dtype) + q[:, :3, None]
    return pts.permute(0, 2, 1)


def dh_matrices(theta, a, d, s_alpha, c_alpha):
    """utils.DH2mat, utils.py:66-77: standard DH 4x4 per joint, (B,J) -> (B,J,4,4)."""
    ct, st = theta.cos(), theta.sin()
    z, o = torch.zeros_like(theta), torch.ones_like(theta)
    rows = [
        [ct, -st * c_alpha, st * s_alpha, a * ct],
        [st, ct * c_alpha, -ct * s_alpha, a * st],
        [z, s_alpha + z, c_alpha + z, d + z],
        [z, z, z, o],
    ]
    return torch.stack([torch.stack(r, dim=2) for r in rows], dim=2)


@dataclass
class DHArm:
    """One serial DH arm: parameters as the reference stores them (float32 tensors, model.py:161-168)."""

    a: torch.Tensor
    d: torch.Tensor
    s_alpha: torch.Tensor
    c_alpha: torch.Tensor
    theta0: torch.Tensor
    mask: Sequence[bool]
    joint_index: Sequence[int]  # which columns of q drive this arm
    base: Optional[torch.Tensor] = None  # (4,4) pre-multiplied base transform (BaxterDualArmFK, model.py:353-360)
    offset: Optional[torch.Tensor] = None  # (3,) translation added to outputs (DualPandaFK, model.py:499-500)
    tool_points: Optional[torch.Tensor] = None  # (T,3) points rigidly attached to the last frame (PandaFK, model.py:445-450)


def fk_dh_arm(q, arm: DHArm):
    """BaxterLeftArmFK.fkine model.py:225-241 / PandaFK.fkine model.py:430-453 for one arm -> (B,M,3)."""
    qa = q[:, list(arm.joint_index)]
    T = dh_matrices(qa + arm.theta0, arm.a, arm.d, arm.s_alpha, arm.c_alpha)
    cur = None
    pts = []
    for i in range(T.shape[1]):
        Ti = T[:, i]
        if cur is None:
            cur = Ti if arm.base is None else arm.base.to(Ti.dtype) @ Ti
        else:
            cur = cur @ Ti
        if arm.mask[i]:
            pts.append(cur[:, :3, 3])
    if arm.tool_points is not None:
        tp = arm.tool_points.to(cur.dtype)
        homog = torch.cat([tp, torch.ones(len(tp), 1, dtype=cur.dtype)], dim=1).T  # (4,T)
        ext = cur @ homog
        pts += [ext[:, :3, j] for j in range(tp.shape[0])]
    out = torch.stack(pts, dim=1)
    if arm.offset is not None:
        out = out + arm.offset
    return out


def fk_dh_multi(q, arms: Sequence[DHArm], dof: int, interleave: bool):
    """Multi-arm DH FK.  interleave=True reproduces BaxterDualArmFK.fkine's point order (model.py:366-383:
    per masked frame, left then right); interleave=False reproduces DualPandaFK.fkine (model.py:486-503:
    all points of arm 0 then all points of arm 1)."""
    q = q.reshape(-1, dof)
    per_arm = [fk_dh_arm(q, a) for a in arms]
    if len(per_arm) == 1:
        return per_arm[0]
    if interleave:
        return torch.stack(per_arm, dim=2).reshape(q.shape[0], -1, 3)
    return torch.cat(per_arm, dim=1)


def fk_se2_base_planar_arm(q, base_keypoints, link_length):
    """cfg-4 composed robot (BASELINE.json configs[3]): an SE(2) rigid base
    (model.py:90-93) carrying a planar revolute chain (model.py:40-48) expressed in the base frame.
    q = (x, y, theta, q1..qK) -> (B, M_base + K, 2)."""
    L = torch.as_tensor(link_length)
    K = L.numel()
    q = q.reshape(-1, 3 + K)
    base_pts = fk_se2_body(q[:, :3], base_keypoints)
    arm_local = fk_planar_chain(q[:, 3:], L)  # (B,K,2) in base frame
    c, s = q[:, 2].cos(), q[:, 2].sin()
    R = torch.stack([torch.stack([c, -s], 1), torch.stack([s, c], 1)], 1)
    arm_world = (R @ arm_local.permute(0, 2, 1) + q[:, :2, None]).permute(0, 2, 1)
    return torch.cat([base_pts, arm_world], dim=1)


# --------------------------------------------------------------------------------------------------
# Scores (diffco/kernel_perceptrons.py, diffco/deprecated/MultiDiffCo.py)
# --------------------------------------------------------------------------------------------------


def fk_joint_tree(q, nodes, n_slots):
    """Link-frame origins of a URDF tree, (B, n_slots, 3) — RigidBody.forward_kinematics, rigid_body.py:86-141, unrolled over
    a parents-first node list instead of recursed: each node = dict(parent, q_index, joint in {'fixed', 'x', 'y', 'z',
    'prismatic'}, sign, axis (3,), rot (3, 3) = Rz(yaw) Ry(pitch) Rx(roll), trans (3,), mimic_mul, mimic_off, out_slot).
    Revolute: R = R_parent rot Rot_axis(sign q'), t = t_parent + R_parent trans (:101-113); prismatic: R = R_parent rot,
    t = t_parent + R_parent (trans + rot axis q') (:114-124); q' = q * multiplier + offset for mimic joints (:93-94)."""
    B = q.shape[0]
    eye = torch.eye(3, dtype=q.dtype).expand(B, 3, 3)
    frames = []
    out = [None] * n_slots
    for nd in nodes:
        Rp, tp = (eye, torch.zeros(B, 3, dtype=q.dtype)) if nd["parent"] < 0 else frames[nd["parent"]]
        rot, trans = nd["rot"].to(q.dtype), nd["trans"].to(q.dtype)
        qv = q[:, nd["q_index"]] * nd["mimic_mul"] + nd["mimic_off"] if nd["q_index"] >= 0 else torch.zeros(B, dtype=q.dtype)
        R = Rp @ rot
        tj = trans.expand(B, 3)
        if nd["joint"] in ("x", "y", "z"):
            a = nd["sign"] * qv
            c, s_, z, o = torch.cos(a), torch.sin(a), torch.zeros_like(a), torch.ones_like(a)
            rows = {"x": [o, z, z, z, c, -s_, z, s_, c], "y": [c, z, s_, z, o, z, -s_, z, c], "z": [c, -s_, z, s_, c, z, z, z, o]}
            R = R @ torch.stack(rows[nd["joint"]], dim=1).reshape(B, 3, 3)
        elif nd["joint"] == "prismatic":
            tj = tj + (rot @ nd["axis"].to(q.dtype)).expand(B, 3) * qv[:, None]
        t = tp + (Rp @ tj[:, :, None])[:, :, 0]
        frames.append((R, t))
        if nd["out_slot"] >= 0:
            out[nd["out_slot"]] = t
    return torch.stack(out, dim=1), frames


def score_original(point, transform, kernel, support_transformed, gains):
    """DiffCo.score_original, kernel_perceptrons.py:362-370 (gains (N,) -> (B,) / 0-dim when B==1);
    legacy MultiDiffCo.score, deprecated/MultiDiffCo.py:118-123 (gains (N,C) -> (B,C))."""
    if point.ndim == 1:
        point = point[None, :]
    if transform is not None:
        point = transform(point)
    return torch.matmul(kernel(point, support_transformed), gains)


def poly_score(point, transform, rbf_kernel, support_transformed, rbf_nodes):
    """DiffCo.poly_score, kernel_perceptrons.py:309-319: cast to rbf_nodes dtype, K @ rbf_nodes[:,None] -> (B,1)."""
    if point.ndim == 1:
        point = point.unsqueeze(0)
    point = point.to(dtype=rbf_nodes.dtype)
    if transform is not None:
        point = transform(point)
    return torch.matmul(rbf_kernel(point, support_transformed), rbf_nodes.unsqueeze(1))


def multi_rbf_score(point, fkine, rbf_kernel, support_fkine, rbf_nodes):
    """legacy MultiDiffCo.rbf_score, deprecated/MultiDiffCo.py:156-170: rbf_nodes (N,C) -> (B,C)."""
    if point.ndim == 1:
        point = point[None, :]
    if fkine is not None:
        point = fkine(point).reshape(len(point), -1)
    return torch.matmul(rbf_kernel(point, support_fkine), rbf_nodes)


def score_and_grad(fn: Callable[[torch.Tensor], torch.Tensor], q: torch.Tensor, grad_out=None):
    """What the optimisers do (optim.py:86-103): s = dist_est(q); (s*go).sum().backward()."""
    q = q.detach().clone().requires_grad_(True)
    s = fn(q)
    go = torch.ones_like(s) if grad_out is None else grad_out.reshape(s.shape)
    (g,) = torch.autograd.grad((s * go).sum(), q)
    return s.detach(), g


# --------------------------------------------------------------------------------------------------
# Training (kernel_perceptrons.py:98-287, deprecated/MultiDiffCo.py:50-154)
# --------------------------------------------------------------------------------------------------


@dataclass
class Perceptron:
    support_points: torch.Tensor
    support_transformed: torch.Tensor
    gains: torch.Tensor
    hypothesis: torch.Tensor
    y: torch.Tensor
    kernel_matrix: torch.Tensor
    support_index: torch.Tensor  # indices into the training set (bit-exact selection gate)
    iterations: int = 0
    rbf_nodes: Optional[torch.Tensor] = None
    trace: List[int] = field(default_factory=list)


def train_perceptron(X, y, kernel, transform=None, beta=1.0, max_iteration=1000, init=None, record_trace=False):
    """DiffCo.train_perceptron + initialize, kernel_perceptrons.py:98-158, 204-220 (max_num_supports=None path).

    Greedy loop: pick the smallest margin (first index on ties); lazily fill its kernel row/column; if the
    margin is non-positive do the gain update, otherwise try to drop the support with the largest positive
    modified margin; stop when neither applies.  ``init`` = (gains, hypothesis, K) for the jump-start path."""
    Xt = X if transform is None else transform(X)
    y = y.reshape(-1)
    n = len(X)
    if init is None:
        gains = torch.zeros(n, dtype=X.dtype)
        K = torch.zeros(n, n, dtype=X.dtype)
        h = torch.zeros(n, dtype=X.dtype)
    else:
        gains, h, K = (t.clone() for t in init)
    it_done = 0
    trace = []
    for it in range(max_iteration):
        it_done = it
        margin = y * h
        mmin, i = torch.min(margin, 0)
        if K[i, i] == 0:
            K[i] = kernel(Xt[i], Xt)
            K[:, i] = K[i]
        if mmin <= 0:
            delta = (beta ** ((1 + y[i]) / 2) * y[i] - h[i]) / K[i, i]
            gains[i] += delta
            h += delta * K[i]
            if record_trace:
                trace.append(int(i))
            continue
        mod = y * (h - gains * torch.diag(K)) * (gains != 0)
        mmax, j = torch.max(mod, 0)
        if mmax > 0 and torch.sum(gains != 0) > 1:
            h -= gains[j] * K[j]
            gains[j] = 0
            if record_trace:
                trace.append(-int(j) - 1)
            continue
        break
    mask = gains != 0
    if mask.sum() < 2:
        mask[torch.where(mask == 0)[0][0]] = True
    idx = torch.where(mask)[0]
    return Perceptron(
        support_points=X[mask], support_transformed=Xt[mask], gains=gains[mask], hypothesis=h[mask], y=y[mask],
        kernel_matrix=K[idx[:, None], idx[None, :]], support_index=idx, iterations=it_done, trace=trace)


def fit_poly(P: Perceptron, rbf_kernel, target="hypo", distance=None):
    """DiffCo.fit_poly, kernel_perceptrons.py:271-287: rbf_nodes = solve(rbf_kernel(S,S), target)."""
    t = {"hypo": P.hypothesis, "label": P.y}.get(target, distance)
    kmat = rbf_kernel(P.support_transformed, P.support_transformed)
    P.rbf_nodes = torch.linalg.solve(kmat, t[:, None]).reshape(-1)
    return P.rbf_nodes


def jump_start(P: Perceptron, X, y, exist_mask, kernel, transform=None):
    """DiffCo.jump_start_initialize, kernel_perceptrons.py:222-269: warm-start state for an update() round.
    Returns (gains, hypothesis, K) over the stacked set X (existing supports where exist_mask is True)."""
    n = len(X)
    novel = X[~exist_mask]
    h = torch.zeros(n, dtype=X.dtype)
    h[exist_mask] = P.hypothesis
    h[~exist_mask] = score_original(novel, transform, kernel, P.support_transformed, P.gains).reshape(-1)
    novel_t = novel if transform is None else transform(novel)
    K = torch.zeros(n, n, dtype=X.dtype)
    e = torch.where(exist_mask)[0]
    v = torch.where(~exist_mask)[0]
    K[e[:, None], e[None, :]] = P.kernel_matrix
    cross = kernel(P.support_transformed, novel_t)
    if cross.ndim == 1:
        cross = cross[None, :]
    K[e[:, None], v[None, :]] = cross
    K[v[:, None], e[None, :]] = cross.T
    gains = torch.zeros(n, dtype=X.dtype)
    gains[exist_mask] = P.gains
    return gains, h, K


def train_multi_perceptron(X, Y, kernel_on_cfg, beta=1.0, max_iteration=1000):
    """legacy MultiDiffCo.train_perceptron + train, deprecated/MultiDiffCo.py:23-83.  Y is (N,C) of +-1;
    ``kernel_on_cfg(x_i, X)`` is the FKKernel-style callable operating on raw configurations.  One kernel
    matrix is shared by all classes; classes are visited in order inside every outer iteration."""
    n, C = Y.shape
    gains = torch.zeros(n, C, dtype=X.dtype)
    H = torch.zeros(n, C, dtype=X.dtype)
    K = torch.zeros(n, n, dtype=X.dtype)
    complete = torch.zeros(C, dtype=torch.bool)
    it_done = 0
    for it in range(max_iteration):
        it_done = it
        margin = Y * H  # evaluated once per outer iteration (deprecated/MultiDiffCo.py:56)
        for c in range(C):
            mmin, i = torch.min(margin[:, c], 0)
            if K[i, i] == 0:
                K[i] = kernel_on_cfg(X[i], X)
                K[:, i] = K[i]
            if mmin <= 0:
                delta = (beta ** ((1 + Y[i, c]) / 2) * Y[i, c] - H[i, c]) / K[i, i]
                gains[i, c] += delta
                H[:, c] += delta * K[i]
                continue
            mod = Y[:, c] * (H[:, c] - gains[:, c] * torch.diag(K)) * (gains[:, c] != 0)
            mmax, j = torch.max(mod, 0)
            if mmax > 0 and torch.sum(gains[:, c] != 0) > 1:
                H[:, c] -= gains[j, c] * K[j]
                gains[j, c] = 0
                continue
            complete[c] = True
        if torch.min(complete):
            break
    keep = torch.sum(gains != 0, dim=1) != 0
    idx = torch.where(keep)[0]
    return Perceptron(
        support_points=X[keep], support_transformed=X[keep], gains=gains[keep], hypothesis=H[keep], y=Y[keep],
        kernel_matrix=K[idx[:, None], idx[None, :]], support_index=idx, iterations=it_done)


def fit_poly_multi(P: Perceptron, rbf_kernel, fkine=None, target="hypo", reg=0.0, distance=None):
    """legacy MultiDiffCo.fit_poly, deprecated/MultiDiffCo.py:125-154: zero the kernel entries that couple a
    support with non-zero gain in class c to one with zero gain in class c (for every c), solve, then zero
    the nodes wherever the gain is zero."""
    X = P.support_points
    if fkine is not None:
        X = fkine(X).reshape(len(X), -1)
    P.support_transformed = X
    t = {"hypo": P.hypothesis, "label": P.y}.get(target, distance)
    kmat = rbf_kernel(X, X).clone()
    C = P.gains.shape[1]
    for c in range(C):
        nz = P.gains[:, c] != 0
        cut = nz[:, None] & (~nz)[None, :]
        kmat[cut | cut.T] = 0
    nodes = torch.linalg.solve(kmat + reg * torch.eye(len(kmat), dtype=kmat.dtype), t)
    nodes[P.gains == 0] = 0
    P.rbf_nodes = nodes
    return nodes


# --------------------------------------------------------------------------------------------------
# Path densification used by the optimisers (utils.py:87-102)
# --------------------------------------------------------------------------------------------------


def dense_path(q, max_step=2.0, max_step_num=None):
    """utils.dense_path, utils.py:87-102: per segment ceil(|dq|/max_step) equally spaced points, plus the last waypoint."""
    if max_step_num is not None:
        tmp = torch.norm(q[1:] - q[:-1], dim=-1).sum().item() / max_step_num
        max_step = max(max_step, tmp)
    out = []
    for i in range(len(q) - 1):
        delta = q[i + 1] - q[i]
        dist = delta.norm()
        steps = int(math.ceil((dist / max_step).item()))
        ir = torch.arange(steps, dtype=q.dtype).reshape(-1, 1)
        out.append(q[i] + ir * delta * max_step / dist)
    out.append(q[-1:])
    return torch.cat(out)


# --------------------------------------------------------------------------------------------------
# Weighted.step (optim.py:686-761): the penalty optimiser's loop, restated
# --------------------------------------------------------------------------------------------------


def weighted_step(p0, fk, score_fn, limits, wrap_fn, *, maxiter, collision_weight, max_move_weight, joint_limit_weight,
                  safety_bias, max_speed, lr, dense_check, mask=None, dif_weight=1.0):
    """optim.py:706-752.  ``score_fn(q) -> (len(q), 1)`` is ``checker.rbf_score``; ``fk`` the robot's feature map;
    Adam with default betas on all waypoints, ``p.grad[~mask] = 0`` before the update, ``robot.wrap`` after it, early
    exit once the constraint loss of the waypoints BEFORE the update is <= 0.5.  Returns (waypoints, steps taken)."""
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=lr)
    steps = 0
    for _ in range(maxiter):
        opt.zero_grad()
        collision = 0
        if collision_weight != 0:
            check_p = dense_path(p, max_step=max_speed) if dense_check else p  # optim.py:709
            collision = torch.clamp(score_fn(check_p) + safety_bias, min=0).mean() * len(p)  # optim.py:710-711
        cp = fk(p)
        seg = (cp[1:] - cp[:-1]).square()
        max_move = torch.clamp(seg.sum(dim=2) - max_speed**2, min=0).sum() if max_move_weight != 0 else 0  # :716-718
        joint_limit = 0
        if joint_limit_weight != 0:
            joint_limit = (torch.clamp(limits[:, 0] - p, min=0) + torch.clamp(p - limits[:, 1], min=0)).sum()  # :721-723
        constraint = collision_weight * collision + max_move_weight * max_move + joint_limit_weight * joint_limit
        loss = dif_weight * seg.sum() + constraint  # optim.py:726-733
        loss.backward()
        if mask is not None:
            p.grad[~mask] = 0.0
        opt.step()
        p.data = wrap_fn(p.data)
        steps += 1
        if float(constraint.detach() if torch.is_tensor(constraint) else constraint) <= 0.5:  # optim.py:747
            break
    return p.detach(), steps


# --------------------------------------------------------------------------------------------------
# High-level checker (collision_checkers.py:163-218, 254-303, 475-503)
# --------------------------------------------------------------------------------------------------


def checker_fit(X, labels01, kernel, transform, verify_ratio, q_verify_fallback=None, init_from=None, exist_mask=None):
    """RBFDiffCo.fit, collision_checkers.py:163-218: labels {0,1} -> +-1; with 0 < verify_ratio < 1 a random verification
    split (torch.randperm, global RNG); train (optionally jump-started, :220-252 -> kernel_perceptrons.py:222-269) with
    max_iteration = len(train set); fit_poly(Polyharmonic(1, 1), 'label'); safety bias = min(|min|, |max|) / 3 of poly_score
    over the verification configurations (:497-503).  Returns (Perceptron, safety_bias, q_verify, labels_verify)."""
    y = 2 * labels01 - 1
    y_verify = None
    if 0 < verify_ratio < 1:
        n_verify = int(verify_ratio * len(X))
        mask = torch.zeros(len(X), dtype=torch.bool)
        mask[torch.randperm(len(X))[:n_verify]] = True
        X_train, q_verify, y_train, y_verify = X[~mask], X[mask], y[~mask], y[mask]
        if exist_mask is not None:
            exist_mask = exist_mask[~mask]
    else:
        X_train, y_train, q_verify = X, y, q_verify_fallback
    init = None
    if init_from is not None:
        init = jump_start(init_from, X_train, y_train, exist_mask, kernel, transform=transform)
    P = train_perceptron(X_train, y_train, kernel, transform=transform, beta=1.0, max_iteration=len(X_train), init=init)
    ph = KernelSpec("polyharmonic", 1.0, 1)
    fit_poly(P, ph, target="label")
    scores = poly_score(q_verify, transform, ph, P.support_transformed, P.rbf_nodes)[:, 0]
    bias = torch.minimum(scores.min().abs(), scores.max().abs()) / 3
    return P, bias, q_verify, y_verify


def checker_rates(P: Perceptron, transform, q, labels_pm1, bias):
    """RBFDiffCo.verify, collision_checkers.py:254-290: (acc, tpr, tnr) of sign(poly_score + bias) — the BIASED rates it returns."""
    ph = KernelSpec("polyharmonic", 1.0, 1)
    pred = 2 * (poly_score(q, transform, ph, P.support_transformed, P.rbf_nodes)[:, 0] + bias > 0).to(q.dtype) - 1
    acc = (pred == labels_pm1).float().mean()
    tpr = (pred[labels_pm1 == 1] == 1).float().mean()
    tnr = (pred[labels_pm1 == -1] == -1).float().mean()
    return acc, tpr, tnr
