"""Robot forward-kinematics feature maps — host-side mirror of the reference's ``diffco/model.py``.

Same class names, constructor arguments and protocol (``dof``, ``limits``, ``fkine(q[, reuse])``, ``wrap(q)``) as
the reference (model.py:9-21 and the classes cited below), so the reference's optimisers and scripts accept these
objects unchanged.  The arithmetic is not done here: every class compiles itself into a ``dc_fk_desc`` (the POD the
CUDA kernels execute, include/diffco_b200.h) and ``fkine`` calls ``dc_fk_forward`` / ``dc_fk_vjp``.  When a robot is
passed as ``transform=`` to ``DiffCo`` the descriptor is fused into the score kernel and FK never touches HBM.

Constants are parameterised exactly as the reference does (float32 tensors, trig of float32 angles), because the
float64 oracle inherits those float32-rounded values through type promotion.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, functional
from ._lib import FkDesc

pi = math.pi


def wrap2pi(theta):
    """utils.wrap2pi, diffco/utils.py:51-52."""
    return (np.pi + theta) % (np.pi * 2) - np.pi


class _FkRunner:
    """Moves queries to the GPU, runs the native FK (or its VJP) and returns results where the query lives."""

    def __init__(self, desc: FkDesc):
        self.desc = desc

    def _to_dev(self, q):
        dev = functional._require_cuda() if not q.is_cuda else q.device
        dt = q.dtype if q.dtype in (torch.float32, torch.float64) else torch.float32
        return q.detach().to(device=dev, dtype=dt)

    def forward(self, q):
        qd = self._to_dev(q)
        return functional.fk_forward(self.desc, qd).to(device=q.device, dtype=qd.dtype)

    def vjp(self, q, g):
        qd = self._to_dev(q)
        return functional.fk_vjp(self.desc, qd, g.to(qd.device)).to(device=q.device, dtype=q.dtype)


class Model:
    """Protocol of diffco/model.py:9-21 plus the native descriptor."""

    dof: int = None
    limits: torch.Tensor = None
    fk_desc: FkDesc = None
    fkine_backup = None

    def _finalize(self):
        self._runner = _FkRunner(self.fk_desc)

    @property
    def n_points(self) -> int:
        return self.fk_desc.n_points

    @property
    def point_dim(self) -> int:
        return self.fk_desc.point_dim

    def fkine(self, q, reuse=False):
        """(…, dof) -> (B, M, d) control points; differentiable w.r.t. q (first order)."""
        if reuse:
            return self.fkine_backup
        q2 = torch.reshape(q, (-1, self.dof))
        if q2.requires_grad and torch.is_grad_enabled():
            x = functional._FkFunction.apply(q2, self._runner)
        else:
            x = self._runner.forward(q2)
        self.fkine_backup = x.reshape(q2.shape[0], self.fk_desc.n_points, self.fk_desc.point_dim)
        return self.fkine_backup

    def polygons(self, q):
        raise NotImplementedError("geometry (FCL) is outside the hot path; see DESIGN.md §7")

    def wrap(self, q):
        return wrap2pi(q)


class ComposedMap(Model):
    """A robot's feature map applied to ``n_repeat`` consecutive blocks of the input and / or with a trailing time
    column passed through: the maps inside the reference's LineFKKernel (kernel.py:145-173; a path segment's two end
    configurations side by side, features concatenated) and TemporalFKKernel (kernel.py:175-202; [q | t] -> [FK(q) | t]).
    Compiles to the robot's own ``dc_fk_desc`` with ``n_repeat`` / ``time_last`` set, so the composite is fused into
    the score kernel like any robot.  ``fkine`` returns (B, F_total, 1)."""

    def __init__(self, robot, n_repeat: int = 1, time_last: bool = False):
        base = robot.fk_desc
        if base is None or base.type == _lib.DC_FK_NONE or base.n_repeat > 1 or base.time_last:
            raise TypeError("ComposedMap wraps a diffco_b200.model robot")
        if n_repeat < 1:
            raise ValueError("n_repeat must be >= 1")
        tl = 1 if time_last else 0
        d = FkDesc.from_buffer_copy(base)
        d.dof = base.dof * n_repeat + tl
        d.n_points = base.n_features * n_repeat + tl
        d.point_dim = 1
        d.n_repeat = n_repeat
        d.time_last = tl
        if d.dof > _lib.DC_MAX_DOF or d.n_points > _lib.DC_MAX_FEATURES:
            raise ValueError(f"composite map too large: {d.dof} columns / {d.n_points} features "
                             f"(limits {_lib.DC_MAX_DOF} / {_lib.DC_MAX_FEATURES})")
        self.robot = robot
        self.fk_desc = d
        self.dof = d.dof
        lim = torch.as_tensor(robot.limits)
        parts = [lim] * n_repeat + ([torch.tensor([[0.0, 1.0]], dtype=lim.dtype)] if tl else [])
        self.limits = torch.cat(parts, dim=0)
        self._finalize()


class RevolutePlanarRobot(Model):
    """model.py:23-76 — planar revolute chain; fkine = cumulative link end points (model.py:40-48)."""

    def __init__(self, link_length, link_width, dof=None, limits=None):
        if limits is None:
            limits = [-np.pi, np.pi]
        if dof is None:
            dof = len(link_length)
        if isinstance(link_length, (int, float)):
            link_length = [link_length] * dof
        if len(limits) == 2 and isinstance(limits[0], (int, float)):
            limits = [limits] * dof
        assert len(limits) == dof and len(link_length) == dof
        if dof > _lib.DC_MAX_LINKS:
            raise ValueError(f"at most {_lib.DC_MAX_LINKS} links are supported")
        self.dof = dof
        self.link_width = link_width
        self.link_length = torch.FloatTensor(link_length)
        self.limits = torch.FloatTensor(limits)
        d = FkDesc()
        d.type, d.dof, d.n_points, d.point_dim, d.n_links = _lib.DC_FK_PLANAR_CHAIN, dof, dof, 2, dof
        for i in range(dof):
            d.link_length[i] = float(self.link_length[i])
        self.fk_desc = d
        self._finalize()


class RigidPlanarBody(Model):
    """model.py:78-117 — SE(2) rigid body; fkine = R(theta) keypoints + (x, y) (model.py:90-93)."""

    def __init__(self, parts, limits=None):
        self.parts = parts
        self.dof = 3
        self.limits = torch.FloatTensor(limits) if limits is not None else torch.FloatTensor(
            [[-10, 10], [-10, 10], [-pi, pi]])
        self.keypoints = torch.FloatTensor([p[1] for p in parts]).T  # 2*M
        m = self.keypoints.shape[1]
        if m > _lib.DC_MAX_KEYPOINTS:
            raise ValueError(f"at most {_lib.DC_MAX_KEYPOINTS} key points are supported")
        d = FkDesc()
        d.type, d.dof, d.n_points, d.point_dim, d.n_keypoints = _lib.DC_FK_SE2_BODY, 3, m, 2, m
        for r in range(2):
            for j in range(m):
                d.keypoints[r][j] = float(self.keypoints[r, j])
        self.fk_desc = d
        self._finalize()

    def wrap(self, q):
        return torch.cat((q[..., :2], wrap2pi(q[..., 2:])), dim=-1)


class RigidBody(Model):
    """model.py:120-174 — SE(3) rigid body; fkine = Rz Ry Rx keypoints + t (model.py:156-159).

    The reference derives default key points from a mesh through trimesh (model.py:148-153); mesh loading is off
    the hot path, so here ``keypoints`` (3 x M) are required and ``body_path`` is only recorded."""

    def __init__(self, body_path=None, keypoints=None, limits=None, transform=None, center=True):
        if keypoints is None:
            raise NotImplementedError("pass keypoints (3 x M); mesh-derived key points need trimesh (off the hot path)")
        self.body_path = body_path
        self.transform = transform
        self.dof = 6
        self.limits = torch.FloatTensor(limits) if limits is not None else torch.FloatTensor(
            [[-10, 10], [-10, 10], [-10, 10], [-pi, pi], [-pi, pi], [-pi, pi]])
        self.keypoints = torch.FloatTensor(keypoints)
        m = self.keypoints.shape[1]
        if self.keypoints.shape[0] != 3 or m > _lib.DC_MAX_KEYPOINTS:
            raise ValueError(f"keypoints must be (3, M<= {_lib.DC_MAX_KEYPOINTS})")
        d = FkDesc()
        d.type, d.dof, d.n_points, d.point_dim, d.n_keypoints = _lib.DC_FK_SE3_BODY, 6, m, 3, m
        for r in range(3):
            for j in range(m):
                d.keypoints[r][j] = float(self.keypoints[r, j])
        self.fk_desc = d
        self._finalize()

    def wrap(self, q):
        return torch.cat((q[..., :3], wrap2pi(q[..., 3:])), dim=-1)


class DHParameters:
    """model.py:161-174."""

    def __init__(self, a=0, alpha=0, d=0, theta=0):
        self.a = torch.FloatTensor(a)
        self.alpha = torch.FloatTensor(alpha)
        self.d = torch.FloatTensor(d)
        self.theta = torch.FloatTensor(theta)
        self.s_alpha = self.alpha.sin()
        self.c_alpha = self.alpha.cos()


def _fill_arm(arm: _lib.DhArm, dh: DHParameters, mask: Sequence[bool], joint_index: Sequence[int], first_slot: int,
              slot_stride: int, base: Optional[torch.Tensor] = None, offset: Optional[torch.Tensor] = None,
              tool_points: Optional[torch.Tensor] = None) -> int:
    """Fill one dc_dh_arm; masked frame k of this arm lands in output slot first_slot + k*slot_stride.
    Returns the number of output points of the arm."""
    n = len(joint_index)
    if n > _lib.DC_MAX_ARM_JOINTS:
        raise ValueError(f"at most {_lib.DC_MAX_ARM_JOINTS} joints per arm")
    arm.n_joints = n
    k = 0
    for i in range(n):
        arm.joint_index[i] = int(joint_index[i])
        arm.a[i], arm.d[i] = float(dh.a[i]), float(dh.d[i])
        arm.s_alpha[i], arm.c_alpha[i] = float(dh.s_alpha[i]), float(dh.c_alpha[i])
        arm.theta0[i] = float(dh.theta[i])
        if mask[i]:
            arm.out_slot[i] = first_slot + k * slot_stride
            k += 1
        else:
            arm.out_slot[i] = -1
    b = torch.eye(4)[:3] if base is None else base[:3]
    for r in range(3):
        for c in range(4):
            arm.base[4 * r + c] = float(b[r, c])
    for r in range(3):
        arm.offset[r] = 0.0 if offset is None else float(offset[r])
    arm.n_tool = 0 if tool_points is None else len(tool_points)
    for t in range(arm.n_tool):
        arm.tool_slot[t] = first_slot + k * slot_stride
        k += 1
        for r in range(3):
            arm.tool[t][r] = float(tool_points[t][r])
    return k


_BAXTER_LIMITS = [[-1.70167993878, 1.70167993878], [-2.147, 1.047], [-3.05417993878, 3.05417993878], [-0.05, 2.618],
                  [-3.059, 3.059], [-1.57079632679, 2.094], [-3.059, 3.059]]
_BAXTER_MASK = [True, False, True, False, True, False, True]


def _baxter_dh():
    # link lengths and DH table of model.py:200-218 (same for both arms, model.py:258-276)
    L = torch.FloatTensor([270.35, 69, 364.35, 69, 374.29, 10, 387.35]) / 1000
    return L, DHParameters(a=[L[1], 0, L[3], 0, L[5], 0, 0], alpha=[-pi / 2, pi / 2, -pi / 2, pi / 2, -pi / 2, pi / 2, 0],
                           d=[L[0], 0, L[2], 0, L[4], 0, L[6]], theta=[0, pi / 2, 0, 0, 0, 0, 0])


class BaxterLeftArmFK(Model):
    """model.py:176-244 — 7-DoF DH chain, control points = origins of frames 0,2,4,6 (fk_mask, model.py:222)."""

    def __init__(self):
        self.limits = torch.FloatTensor(_BAXTER_LIMITS)
        self.L, self.dhparams = _baxter_dh()
        self.c_alpha, self.s_alpha = self.dhparams.alpha.cos(), self.dhparams.alpha.sin()
        self.dof = 7
        self.fk_mask = list(_BAXTER_MASK)
        d = FkDesc()
        d.type, d.dof, d.point_dim, d.n_arms = _lib.DC_FK_DH_ARMS, 7, 3, 1
        d.n_points = _fill_arm(d.arms[0], self.dhparams, self.fk_mask, range(7), 0, 1)
        self.fk_desc = d
        self._finalize()


class BaxterRightArmFK(BaxterLeftArmFK):
    """model.py:246-311 — identical kinematic table to the left arm in the reference."""


BaxterFK = BaxterLeftArmFK  # model.py:388


class BaxterDualArmFK(Model):
    """model.py:313-386 — both arms behind base transforms (model.py:353-360); output order per masked frame is
    (left, right) (model.py:366-383)."""

    def __init__(self):
        self.limits = torch.FloatTensor(_BAXTER_LIMITS).repeat(2, 1)
        self.L, self.left_dhparams = _baxter_dh()
        _, self.right_dhparams = _baxter_dh()
        offsets = torch.FloatTensor([278, 64, 1104]) / 1000
        # the reference builds the bases in float32 (utils.rotz of a float32 angle, model.py:353-360)
        lt = torch.tensor([-pi / 4])
        rt = torch.tensor([-3 * pi / 4])
        left_base = torch.zeros(4, 4)
        left_base[:3, :3] = torch.tensor([[lt.cos().item(), -lt.sin().item(), 0], [lt.sin().item(), lt.cos().item(), 0], [0, 0, 1]])
        left_base[:, 3] = torch.tensor([offsets[0], -offsets[1], offsets[2], 1])
        right_base = torch.zeros(4, 4)
        right_base[:3, :3] = torch.tensor([[rt.cos().item(), -rt.sin().item(), 0], [rt.sin().item(), rt.cos().item(), 0], [0, 0, 1]])
        right_base[:, 3] = torch.tensor([-offsets[0], -offsets[1], offsets[2], 1])
        self.arm_bases = torch.stack([left_base, right_base])[None, :]
        self.dof = 14
        self.fk_mask = list(_BAXTER_MASK)
        d = FkDesc()
        d.type, d.dof, d.point_dim, d.n_arms = _lib.DC_FK_DH_ARMS, 14, 3, 2
        nl = _fill_arm(d.arms[0], self.left_dhparams, self.fk_mask, range(0, 7), 0, 2, base=left_base)
        nr = _fill_arm(d.arms[1], self.right_dhparams, self.fk_mask, range(7, 14), 1, 2, base=right_base)
        d.n_points = nl + nr
        self.fk_desc = d
        self._finalize()


_PANDA_LIMITS = [[-2.8973, 2.8973], [-1.7628, 1.7628], [-2.8973, 2.8973], [-3.0718, -0.0698], [-2.8973, 2.8973],
                 [-0.0175, 3.7525], [-2.8973, 2.8973]]


def _panda_dh():
    L = torch.FloatTensor([0.3330, 0.3160, 0.0825, 0.3840, 0.0880, 0.1070 * 2])
    return L, DHParameters(a=[0, 0, L[2], -L[2], 0, L[4], 0], alpha=[-pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0],
                           d=[L[0], 0, L[1], 0, L[3], 0, L[5]], theta=[0, 0, 0, 0, 0, 0, 0])


class PandaFK(Model):
    """model.py:390-453 — Franka Panda; frames 0,2,3,4,6 plus two finger points at +-0.5 d7 along the last frame's y
    (model.py:445-450) -> 7 control points.  ``finger_points=False`` gives the 5-point map of the older twin
    diffco/robot_fkine.py:428-444."""

    def __init__(self, finger_points: bool = True):
        self.limits = torch.FloatTensor(_PANDA_LIMITS)
        self.L, self.dhparams = _panda_dh()
        self.c_alpha, self.s_alpha = self.dhparams.alpha.cos(), self.dhparams.alpha.sin()
        self.dof = 7
        self.fk_mask = [True, False, True, True, True, False, True]
        d = FkDesc()
        d.type, d.dof, d.point_dim, d.n_arms = _lib.DC_FK_DH_ARMS, 7, 3, 1
        d.n_points = _fill_arm(d.arms[0], self.dhparams, self.fk_mask, range(7), 0, 1, tool_points=self._tools(finger_points))
        self.fk_desc = d
        self._finalize()

    def _tools(self, finger_points):
        if not finger_points:
            return None
        half = 0.5 * float(self.dhparams.d[-1])
        return torch.tensor([[0.0, half, 0.0], [0.0, -half, 0.0]])


class DualPandaFK(Model):
    """model.py:456-503 — two Pandas with interleaved joints (odd columns -> first arm, even -> second) and base
    offsets added to the outputs (model.py:490-500); output = first arm's 7 points then second arm's."""

    def __init__(self):
        self.limits = torch.FloatTensor([l for l in _PANDA_LIMITS for _ in range(2)])
        self.left_panda, self.right_panda = PandaFK(), PandaFK()
        right_base = torch.FloatTensor([0.0, 0.0, 0.0])
        left_base = torch.FloatTensor([0.0, 0.84, 0.0])
        self.bases = torch.stack([left_base, right_base], dim=0)
        self.dhparams = self.left_panda.dhparams
        self.dof = 14
        tools = self.left_panda._tools(True)
        mask = self.left_panda.fk_mask
        d = FkDesc()
        d.type, d.dof, d.point_dim, d.n_arms = _lib.DC_FK_DH_ARMS, 14, 3, 2
        # model.py:492-500: left_q = odd columns, offset bases[0]; right_q = even columns, offset bases[1]
        n0 = _fill_arm(d.arms[0], self.dhparams, mask, [1, 3, 5, 7, 9, 11, 13], 0, 1, offset=self.bases[0], tool_points=tools)
        n1 = _fill_arm(d.arms[1], self.dhparams, mask, [0, 2, 4, 6, 8, 10, 12], n0, 1, offset=self.bases[1], tool_points=tools)
        d.n_points = n0 + n1
        self.fk_desc = d
        self._finalize()


class SE2BasePlanarArm(Model):
    """SE(2) mobile base (x, y, theta) carrying a planar revolute chain — BASELINE.json configs[3].  The reference has
    no such class; it is composed from RigidPlanarBody.fkine (model.py:90-93) for the base key
    points and RevolutePlanarRobot.fkine (model.py:40-48) evaluated in the base frame for the arm."""

    def __init__(self, base_keypoints, link_length, limits=None):
        self.keypoints = torch.FloatTensor(base_keypoints)  # 2 x M
        self.link_length = torch.FloatTensor(link_length)
        mb, k = self.keypoints.shape[1], len(self.link_length)
        self.dof = 3 + k
        if limits is None:
            limits = [[-10, 10], [-10, 10], [-pi, pi]] + [[-pi, pi]] * k
        self.limits = torch.FloatTensor(limits)
        d = FkDesc()
        d.type, d.dof, d.n_points, d.point_dim = _lib.DC_FK_SE2_BASE_PLANAR_ARM, self.dof, mb + k, 2
        d.n_keypoints, d.n_links = mb, k
        for r in range(2):
            for j in range(mb):
                d.keypoints[r][j] = float(self.keypoints[r, j])
        for i in range(k):
            d.link_length[i] = float(self.link_length[i])
        self.fk_desc = d
        self._finalize()

    def wrap(self, q):
        return torch.cat((q[..., :2], wrap2pi(q[..., 2:])), dim=-1)
