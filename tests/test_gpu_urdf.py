"""GPU parity (-m gpu) of the URDF joint program (DC_FK_JOINT_TREE): the device forward kinematics, its
analytic J^T product, the all-links frames and the fused score + gradient, against the reference's RigidBody recursion
(tests/golden/urdf.npz, float32 in the reference) and the float64 oracle restatement; then ForwardKinematicsDiffCo built
straight from a URDF path."""
import os

import numpy as np
import pytest
import torch

from tests import problems as P

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)
T64 = lambda a: torch.from_numpy(np.asarray(a)).double()
URDF = {"arm7": "arm7_gripper.urdf", "torso": "torso_two_arms.urdf"}
BASE = {"arm7": None, "torso": torch.tensor([[0.0, -1.0, 0.0, 0.3], [1.0, 0.0, 0.0, -0.2], [0.0, 0.0, 1.0, 0.1], [0, 0, 0, 1.0]]),
        "panda": None}


def rel(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "urdf.npz"))


def _robot(name, g):
    from diffco_b200 import _lib
    from diffco_b200 import model as M
    from diffco_b200.collision_interfaces import URDFRobot

    if name in URDF:
        return URDFRobot(os.path.join(HERE, "data", URDF[name]), base_transform=BASE[name])
    # the reference's Panda description does not travel: rebuild the robot around the descriptor the golden stores
    robot = M.Model()
    robot.fk_desc = _lib.FkDesc.from_buffer_copy(g[name + "_desc"].tobytes())
    robot.dof = robot.fk_desc.dof
    robot.limits = T64(g[name + "_limits"]).float()
    robot._finalize()
    return robot


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", ["arm7", "torso", "panda"])
def test_tree_fk_forward_and_vjp(name, dtype, g, cuda_device):
    from diffco_b200 import functional as Fn
    from oracle import diffco_oracle as O

    robot = _robot(name, g)
    desc = robot.fk_desc
    q = T64(g[name + "_q"]).to(device=cuda_device, dtype=dtype)
    x = Fn.fk_forward(desc, q).reshape(len(q), -1, 3)
    assert rel(x.transpose(1, 2), g[name + "_x"]) <= 3e-6  # the reference computes in float32
    gx = T64(g[name + "_gx"]).transpose(1, 2).contiguous()  # golden layout (B, 3, L) -> ours (B, L, 3)
    gq = Fn.fk_vjp(desc, q, gx.to(device=cuda_device, dtype=dtype))
    assert rel(gq, g[name + "_gq"]) <= 1e-5
    # float64 oracle on the same joint program: the arithmetic gate
    qo = T64(g[name + "_q"]).requires_grad_(True)
    xo, _ = O.fk_joint_tree(qo, P.tree_nodes_from_desc(desc), desc.n_points)
    (xo * gx).sum().backward()
    assert rel(x, xo) <= (2e-7 if dtype == torch.float32 else 1e-13)
    assert rel(gq, qo.grad) <= (2e-6 if dtype == torch.float32 else 1e-12)
    # Model.fkine protocol: CPU query in, differentiable
    qv = T64(g[name + "_q"]).to(dtype).requires_grad_(True)
    pts = robot.fkine(qv)
    assert pts.shape == (len(q), desc.n_points, 3) and pts.device.type == "cpu"
    (pts * gx.to(dtype)).sum().backward()
    assert rel(qv.grad, qo.grad) <= (2e-6 if dtype == torch.float32 else 1e-12)


@pytest.mark.parametrize("name", ["arm7", "torso"])
def test_all_link_frames_match_reference(name, g, cuda_device):
    robot = _robot(name, g)
    q = T64(g[name + "_q"]).float().to(cuda_device)
    poses = robot.compute_forward_kinematics_all_links(q)
    assert list(poses.keys()) == list(g[name + "_nodes"]) and set(poses) == set(g[name + "_links"])
    trans = torch.stack([poses[k][0][0] for k in g[name + "_links"]], 1)
    rot = torch.stack([poses[k][0][1] for k in g[name + "_links"]], 1)
    assert trans.is_cuda and rel(trans, g[name + "_trans"]) <= 3e-6 and rel(rot, g[name + "_rot"]) <= 3e-6
    if name == "arm7":  # collision-geometry poses: link pose composed with the <collision> origin (rigid_body.py:128-129)
        cp = robot.compute_forward_kinematics_all_links(q, return_collision=True)
        assert len(cp["l3"]) == 2 and len(cp["l2"]) == 0 and len(cp["tool"]) == 1
        t_l3, r_l3 = poses["l3"][0]
        want = t_l3 + (r_l3 @ torch.tensor([0.0, 0.1, 0.0], device=cuda_device))
        assert rel(cp["l3"][1][0], want) <= 1e-6 and rel(cp["l3"][1][1], r_l3) <= 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name,batch", [("arm7", 300), ("panda", 5000)])
def test_fused_score_and_gradient_on_urdf_features(name, batch, dtype, g, cuda_device):
    """DiffCo(transform=robot.fkine): the joint program runs inside the score kernels (lane-split for the small batch,
    thread-per-query / lane-split for the large one); oracle: same model in float64 through autograd."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K
    from oracle import diffco_oracle as O

    robot = _robot(name, g)
    desc = robot.fk_desc
    gen = torch.Generator().manual_seed(3)
    lim = robot.limits.double()
    S = torch.rand(400, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    Q = torch.rand(batch, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    W = torch.randn(400, generator=gen, dtype=torch.float64)
    S, Q, W = (t.to(dtype).double() for t in (S, Q, W))
    dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine)
    dc.support_points = S.to(device=cuda_device, dtype=dtype)
    dc.support_transformed = robot.fkine(dc.support_points)
    dc.gains = W.to(device=cuda_device, dtype=dtype)
    s, gq = dc.score_and_grad(Q.to(device=cuda_device, dtype=dtype))
    nodes = P.tree_nodes_from_desc(desc)
    fk = lambda q: O.fk_joint_tree(q, nodes, desc.n_points)[0]
    St = fk(S)
    so, go = O.score_and_grad(lambda q: O.score_original(q, fk, O.KernelSpec("rq", 10.0, 2), St, W), Q)
    tol = 1e-5 if dtype == torch.float32 else 1e-10
    assert rel(s.reshape(-1), so.reshape(-1)) <= tol and rel(gq, go) <= tol


def test_fk_checker_from_urdf_path(cuda_device):
    """ForwardKinematicsDiffCo(robot=<path to a URDF>) (collision_checkers.py:52-56,318-372) with an injected ground truth:
    a ball around a point the tool can reach."""
    from diffco_b200 import ForwardKinematicsDiffCo

    torch.manual_seed(0)
    path = os.path.join(HERE, "data", URDF["arm7"])
    probe = {}

    def gt(q):
        tool = probe["robot"].fkine(q)[:, 5, :]  # flange origin
        return ((tool - torch.tensor([0.3, 0.2, 0.5], device=tool.device)).norm(dim=1) < 0.35).to(q.dtype)

    checker = ForwardKinematicsDiffCo(robot=path, gt_check_func=gt, device=cuda_device)
    probe["robot"] = checker.robot.model
    assert checker.robot._n_dofs == 8 and checker.robot.model.unique_position_link_names[5] == "flange"
    acc, tpr, tnr = checker.fit(num_samples=3000, verify_ratio=0.1)
    assert acc > 0.85
    q = checker.robot.rand_configs(2000)
    score = checker.collision_score(q).reshape(-1)
    agree = ((score > 0) == (gt(q) > 0)).float().mean()
    assert float(agree) > 0.85


def test_multi_robot_feature_map_is_the_concatenation(g, cuda_device):
    """MultiURDFRobot: the merged joint program equals the robots' own maps side by side (bit for bit), its J^T product the
    per-robot products, and the per-robot frame dictionaries those of the single robots."""
    from diffco_b200 import functional as Fn
    from diffco_b200.collision_interfaces import MultiURDFRobot, URDFRobot

    a = URDFRobot(os.path.join(HERE, "data", URDF["arm7"]), name="arm")
    b = URDFRobot(os.path.join(HERE, "data", URDF["torso"]), name="torso", base_transform=BASE["torso"])
    multi = MultiURDFRobot(urdf_robots=[a, b])
    assert multi.dof == 15 and multi.fk_desc.n_points == 16
    q = torch.cat([T64(g["arm7_q"]), T64(g["torso_q"])], dim=1).to(cuda_device)
    x = multi.fkine(q)
    xa, xb = a.fkine(q[:, :8]), b.fkine(q[:, 8:])
    assert x.shape == (16, 16, 3) and torch.equal(x, torch.cat([xa, xb], dim=1))
    gx = torch.randn(16, 16, 3, generator=torch.Generator().manual_seed(8), dtype=torch.float64).to(cuda_device)
    gq = Fn.fk_vjp(multi.fk_desc, q, gx)
    want = torch.cat([Fn.fk_vjp(a.fk_desc, q[:, :8].contiguous(), gx[:, :8].contiguous()),
                      Fn.fk_vjp(b.fk_desc, q[:, 8:].contiguous(), gx[:, 8:].contiguous())], dim=1)
    assert torch.equal(gq, want)
    dicts = multi.compute_forward_kinematics_all_links(q.float())
    assert len(dicts) == 2 and set(dicts[0]) == set(g["arm7_links"]) and set(dicts[1]) == set(g["torso_links"])
    assert rel(torch.stack([dicts[1][k][0][0] for k in g["torso_links"]], 1), g["torso_trans"]) <= 3e-6
