"""Deterministic synthetic problems shared by the parity tests, smoke() and bench.py.

Everything is generated on the CPU in float64 from explicit seeds and cast afterwards, so the oracle and the CUDA
path always see identical inputs.  Robots are built twice — the product's descriptor classes (diffco_b200.model) and
the matching oracle closures (oracle/diffco_oracle.py) parameterised from the very same constants.
"""
from __future__ import annotations

import torch

from oracle import diffco_oracle as O


def _arm_from_desc(desc_arm, n_points_stride=None):
    """Oracle DHArm from a product dc_dh_arm (so both sides share the float32-rounded constants)."""
    J = desc_arm.n_joints
    f = lambda arr, n=J: torch.tensor([arr[i] for i in range(n)], dtype=torch.float64)
    base = torch.eye(4, dtype=torch.float64)
    base[:3, :] = torch.tensor([desc_arm.base[i] for i in range(12)], dtype=torch.float64).reshape(3, 4)
    tools = None
    if desc_arm.n_tool:
        tools = torch.tensor([[desc_arm.tool[t][r] for r in range(3)] for t in range(desc_arm.n_tool)], dtype=torch.float64)
    return O.DHArm(a=f(desc_arm.a), d=f(desc_arm.d), s_alpha=f(desc_arm.s_alpha), c_alpha=f(desc_arm.c_alpha),
                   theta0=f(desc_arm.theta0), mask=[desc_arm.out_slot[i] >= 0 for i in range(J)],
                   joint_index=[desc_arm.joint_index[i] for i in range(J)], base=base,
                   offset=torch.tensor([desc_arm.offset[i] for i in range(3)], dtype=torch.float64), tool_points=tools)


def oracle_fk(robot):
    """Oracle feature map (q (B,D) float64 -> (B,M,d)) equivalent to a diffco_b200.model robot."""
    from diffco_b200 import _lib

    d = robot.fk_desc
    if d.type == _lib.DC_FK_PLANAR_CHAIN:
        L = torch.tensor([d.link_length[i] for i in range(d.n_links)], dtype=torch.float64)
        return lambda q: O.fk_planar_chain(q, L)
    if d.type == _lib.DC_FK_SE2_BODY:
        kp = torch.tensor([[d.keypoints[r][j] for j in range(d.n_keypoints)] for r in range(2)], dtype=torch.float64)
        return lambda q: O.fk_se2_body(q, kp)
    if d.type == _lib.DC_FK_SE3_BODY:
        kp = torch.tensor([[d.keypoints[r][j] for j in range(d.n_keypoints)] for r in range(3)], dtype=torch.float64)
        return lambda q: O.fk_se3_body(q, kp)
    if d.type == _lib.DC_FK_SE2_BASE_PLANAR_ARM:
        kp = torch.tensor([[d.keypoints[r][j] for j in range(d.n_keypoints)] for r in range(2)], dtype=torch.float64)
        L = torch.tensor([d.link_length[i] for i in range(d.n_links)], dtype=torch.float64)
        return lambda q: O.fk_se2_base_planar_arm(q, kp, L)
    if d.type == _lib.DC_FK_DH_ARMS:
        arms = [_arm_from_desc(d.arms[a]) for a in range(d.n_arms)]
        # interleaved output (BaxterDualArmFK) <=> arm 1's first slot is 1
        interleave = d.n_arms == 2 and min(s for s in d.arms[1].out_slot[: d.arms[1].n_joints] if s >= 0) == 1
        return lambda q: O.fk_dh_multi(q, arms, d.dof, interleave)
    raise ValueError(d.type)


def make_robot(name):
    from diffco_b200 import model as M

    if name == "planar2":
        return M.RevolutePlanarRobot(1.0, 0.3, dof=2)
    if name == "planar3":
        return M.RevolutePlanarRobot([1.0, 0.8, 0.6], 0.3)
    if name == "planar7":
        return M.RevolutePlanarRobot(1.0, 0.3, dof=7)
    if name == "planar15":  # F = 30: the widest map the tensor-core image holds
        return M.RevolutePlanarRobot(0.5, 0.1, dof=15)
    if name == "se2":
        return M.RigidPlanarBody([("box", (0.5, 0.2), (1, 1)), ("box", (-0.4, 0.3), (1, 1)), ("box", (0.1, -0.6), (1, 1)),
                                  ("box", (-0.3, -0.2), (1, 1)), ("box", (0.7, 0.7), (1, 1))])
    if name == "se3":
        c = [[sx * 0.5, sy * 0.3, sz * 0.2] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        return M.RigidBody(keypoints=torch.tensor(c).T.tolist())
    if name == "baxter":
        return M.BaxterLeftArmFK()
    if name == "baxter_dual":
        return M.BaxterDualArmFK()
    if name == "panda":
        return M.PandaFK()
    if name == "panda5":
        return M.PandaFK(finger_points=False)
    if name == "dual_panda":
        return M.DualPandaFK()
    if name == "se2arm":
        return M.SE2BasePlanarArm([[0.5, -0.5, -0.5, 0.5], [0.3, 0.3, -0.3, -0.3]], [1.0, 1.0, 1.0])
    raise KeyError(name)


ROBOTS = ["planar2", "planar3", "planar7", "se2", "se3", "baxter", "baxter_dual", "panda", "panda5", "dual_panda", "se2arm"]


def sample_configs(robot, n, gen):
    lim = robot.limits.double()
    return torch.rand(n, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]


def synthetic_model(robot_name, n_sv, n_class, seed):
    """(robot, support configs (N,D), weights (N,C)) — weights ~ N(0,1) with a per-class sparsity pattern when C>1
    (rows zero where that class's gain is zero, like a trained MultiDiffCo)."""
    gen = torch.Generator().manual_seed(seed)
    robot = make_robot(robot_name)
    S = sample_configs(robot, n_sv, gen)
    W = torch.randn(n_sv, n_class, generator=gen, dtype=torch.float64)
    if n_class > 1:
        W = W * (torch.rand(n_sv, n_class, generator=gen) < 0.6)
    return robot, S, W


def circle_labels(robot, q, circles=(((3.0, 2.0), 2.0), ((-2.0, 3.0), 0.8))):
    """+1 when any control point lies inside any circle, else -1 (planar robots; SURVEY.md §8d cfg-2 'trained')."""
    pts = oracle_fk(robot)(q.double())
    hit = torch.zeros(len(q), dtype=torch.bool)
    for (cx, cy), r in circles:
        hit |= ((pts - torch.tensor([cx, cy], dtype=torch.float64)).norm(dim=2) < r).any(dim=1)
    return hit.double() * 2 - 1


def rel_to_max(a, b):
    """max|a-b| / max|b| — the parity metric of BASELINE.md §3."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def tree_nodes_from_desc(desc):
    """dc_fk_desc.tree -> the node list oracle.fk_joint_tree takes."""
    kinds = {0: "fixed", 1: "x", 2: "y", 3: "z", 4: "prismatic"}
    nodes = []
    for i in range(desc.n_nodes):
        nd = desc.tree[i]
        nodes.append({"parent": nd.parent, "q_index": nd.q_index, "joint": kinds[nd.joint], "sign": nd.axis[0],
                      "axis": torch.tensor(list(nd.axis), dtype=torch.float64), "out_slot": nd.out_slot,
                      "rot": torch.tensor(list(nd.rot), dtype=torch.float64).reshape(3, 3),
                      "trans": torch.tensor(list(nd.trans), dtype=torch.float64), "mimic_mul": nd.mimic_mul,
                      "mimic_off": nd.mimic_off})
    return nodes
