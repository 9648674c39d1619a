"""GPU parity (-m gpu) of the composite kernels of kernel.py:145-202 — TemporalFKKernel ([q | t] rows, DC_K_RQ_TEMPORAL on
[FK(q) | t] features), LineFKKernel ([q_a | q_b] rows, FK repeated on both halves) and LineKernel — against the reference's
own outputs (tests/golden/line_temporal.npz, oracle/make_golden.py::gen_line_temporal): kernel matrices, the fused
score + analytic gradient through DiffCo (lane-split kernel with the composite feature map), and training on segments."""
import os

import numpy as np
import pytest
import torch

from tests import problems as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
T64 = lambda a: torch.from_numpy(np.asarray(a)).double()


def rel(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "line_temporal.npz"))


def _kernels(name, g, robot):
    from diffco_b200 import kernel as K

    if name == "line":
        return K.LineFKKernel(robot.fkine, K.RQKernel(10.0)), "l_x", "l_s", "l"
    gx, px, gt, pt, al = g[name + "_params"]
    return K.TemporalFKKernel(robot.fkine, K.RQKernel(gx, int(px)), K.RQKernel(gt, int(pt)), alpha=al), "t_x", "t_s", name


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", ["t_a", "t_b", "line"])
def test_kernel_matrix_and_fused_score_gradient(name, dtype, g, cuda_device):
    from diffco_b200 import DiffCo

    robot = P.make_robot("planar7")
    kfun, kx, ks, tag = _kernels(name, g, robot)
    tol = 3e-6 if dtype == torch.float32 else 1e-11
    x, s, w = T64(g[kx]).to(dtype), T64(g[ks]).to(dtype), T64(g["w"]).to(dtype)
    k = kfun(x.to(cuda_device), s.to(cuda_device))
    assert rel(k, g[tag + "_K"]) <= tol
    k1 = kfun(x[0].to(cuda_device), s.to(cuda_device))  # single row: squeezed like RQKernel (kernel.py:26-27)
    assert rel(k1, g[tag + "_K_single"]) <= tol

    dc = DiffCo(kernel_func=kfun)
    dc.support_points = s.to(cuda_device)
    dc.support_transformed = dc._shape_features(dc._features(dc.support_points))
    dc.gains = w.to(cuda_device)
    want = g[tag + "_K"] @ g["w"]
    xv = x.clone().requires_grad_(True)  # CPU rows in, like the reference's optimisers
    sc = dc.score(xv)
    assert rel(sc.reshape(-1), want) <= tol
    sc.sum().backward()
    assert rel(xv.grad, g[tag + "_grad"]) <= (1e-5 if dtype == torch.float32 else 1e-10)
    s2, g2 = dc.score_and_grad(x.to(cuda_device))
    assert rel(s2.reshape(-1), want) <= tol and rel(g2, g[tag + "_grad"]) <= (1e-5 if dtype == torch.float32 else 1e-10)


def test_line_kernel_matrix(g, cuda_device):
    from diffco_b200 import kernel as K

    lk = K.LineKernel(K.RQKernel(2.0))
    k = lk(T64(g["l_x"]).to(cuda_device), T64(g["l_s"]).to(cuda_device))
    assert rel(k, g["lk_K"]) <= 1e-12


def test_large_batch_of_segments_matches_float64(cuda_device):
    """A line-query batch large enough for every launch-shape branch (B = 5000 segments, N = 600 supports), float32
    against the float64 evaluation of the same model."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K

    robot = P.make_robot("planar3")
    gen = torch.Generator().manual_seed(5)
    S = torch.cat([P.sample_configs(robot, 600, gen), P.sample_configs(robot, 600, gen)], 1)
    Q = torch.cat([P.sample_configs(robot, 5000, gen), P.sample_configs(robot, 5000, gen)], 1)
    W = torch.randn(600, generator=gen, dtype=torch.float64)
    out = {}
    for dtype in (torch.float64, torch.float32):
        dc = DiffCo(kernel_func=K.LineFKKernel(robot.fkine, K.RQKernel(10.0)))
        dc.support_points = S.to(device=cuda_device, dtype=dtype)
        dc.support_transformed = dc._shape_features(dc._features(dc.support_points))
        dc.gains = W.to(device=cuda_device, dtype=dtype)
        out[dtype] = dc.score_and_grad(Q.to(device=cuda_device, dtype=dtype))
    s64, g64 = out[torch.float64]
    s32, g32 = out[torch.float32]
    assert rel(s32, s64.cpu().numpy()) <= 1e-5 and rel(g32, g64.cpu().numpy()) <= 1e-5


def test_training_on_space_time_points(cuda_device):
    """DiffCo.train with the temporal kernel: a moving obstacle's labels are separated (every training point classified
    correctly, the perceptron's stopping rule) and the hypothesis equals K @ gains."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K

    robot = P.make_robot("planar3")
    gen = torch.Generator().manual_seed(9)
    q = P.sample_configs(robot, 800, gen)
    t = torch.rand(800, 1, generator=gen, dtype=torch.float64)
    X = torch.cat([q, t], 1)
    tip = robot.fkine(q.float())[:, -1, :].double().cpu()
    centre = torch.stack([1.5 - 3.0 * t[:, 0], torch.full((800,), 1.0, dtype=torch.float64)], 1)  # obstacle sweeps in x
    y = torch.where((tip - centre).norm(dim=1) < 0.8, 1.0, -1.0)
    kfun = K.TemporalFKKernel(robot.fkine, K.RQKernel(10.0), K.RQKernel(10.0), alpha=1.0)
    dc = DiffCo(kernel_func=kfun, beta=1.0)
    dc.train(X.to(cuda_device), y.to(cuda_device), max_iteration=5000)
    assert dc.gains is not None and len(dc.support_points) > 0
    sc = dc.score(X.to(cuda_device)).reshape(-1).cpu()
    assert bool(((sc > 0) == (y > 0)).all())
    Kmat = kfun(X.to(cuda_device), dc.support_points)
    assert rel(sc, (Kmat.cpu() @ dc.gains.reshape(-1).cpu()).numpy()) <= 1e-9


@pytest.mark.parametrize("rname", ["baxter", "se2", "urdf_torso"])
def test_composite_maps_on_other_robots_match_the_oracle(rname, cuda_device):
    """LineFKKernel / TemporalFKKernel wrap ANY robot's map (DH arms, rigid bodies, URDF trees): fused score + gradient in
    float64 against the oracle's statement of kernel.py:145-202 on the oracle's own FK of that robot."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K
    from oracle import diffco_oracle as O

    if rname == "urdf_torso":
        from diffco_b200.collision_interfaces import URDFRobot

        robot = URDFRobot(os.path.join(os.path.dirname(__file__), "data", "torso_two_arms.urdf"))
        nodes = P.tree_nodes_from_desc(robot.fk_desc)
        fk = lambda q: O.fk_joint_tree(q, nodes, robot.fk_desc.n_points)[0]
    else:
        robot = P.make_robot(rname)
        fk = P.oracle_fk(robot)
    gen = torch.Generator().manual_seed(21)
    lim = robot.limits.double()
    draw = lambda n: torch.rand(n, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    w = torch.randn(40, generator=gen, dtype=torch.float64)
    # segments [q_a | q_b]
    S, Q = torch.cat([draw(40), draw(40)], 1), torch.cat([draw(25), draw(25)], 1)
    Q[3] = S[5]
    want = lambda q: O.line_fk_kernel(q, S, fk, 10.0) @ w
    s_ref, g_ref = O.score_and_grad(want, Q)
    dc = DiffCo(kernel_func=K.LineFKKernel(robot.fkine, K.RQKernel(10.0)))
    dc.support_points = S.to(cuda_device)
    dc.support_transformed = dc._shape_features(dc._features(dc.support_points))
    dc.gains = w.to(cuda_device)
    s, g = dc.score_and_grad(Q.to(cuda_device))
    assert rel(s.reshape(-1), s_ref.reshape(-1).numpy()) <= 1e-10 and rel(g, g_ref.numpy()) <= 1e-9
    # space-time points [q | t]
    St, Qt = torch.cat([draw(40), torch.rand(40, 1, generator=gen, dtype=torch.float64)], 1), torch.cat(
        [draw(25), torch.rand(25, 1, generator=gen, dtype=torch.float64)], 1)
    want_t = lambda q: O.temporal_fk_kernel(q, St, fk, 10.0, 2, 4.0, 1, 0.7) @ w
    s_ref, g_ref = O.score_and_grad(want_t, Qt)
    dct = DiffCo(kernel_func=K.TemporalFKKernel(robot.fkine, K.RQKernel(10.0), K.RQKernel(4.0, 1), alpha=0.7))
    dct.support_points = St.to(cuda_device)
    dct.support_transformed = dct._shape_features(dct._features(dct.support_points))
    dct.gains = w.to(cuda_device)
    s, g = dct.score_and_grad(Qt.to(cuda_device))
    assert rel(s.reshape(-1), s_ref.reshape(-1).numpy()) <= 1e-10 and rel(g, g_ref.numpy()) <= 1e-9
