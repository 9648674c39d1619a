"""CPU tests: pin the oracle (oracle/diffco_oracle.py, oracle/analytic_np.py) against the golden fixtures that
oracle/make_golden.py produced by running the unmodified reference, and — when /root/reference is present —
against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import analytic_np as A
from oracle import diffco_oracle as O
from oracle import ref_loader
from tests import problems as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")
T64 = lambda a: torch.from_numpy(np.asarray(a)).double()


def load(name):
    return np.load(os.path.join(GOLD, name))


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-300)
    assert np.abs(a - b).max() / scale <= tol, np.abs(a - b).max() / scale


KERNEL_CASES = {
    "rq_g10_p2": O.KernelSpec("rq", 10.0, 2), "rq_g3_p3": O.KernelSpec("rq", 3.0, 3), "rq_g1_p1": O.KernelSpec("rq", 1.0, 1),
    "ph_k1_e1": O.KernelSpec("polyharmonic", 1.0, 1), "ph_k3_e05": O.KernelSpec("polyharmonic", 0.5, 3),
    "ph_k2_e1": O.KernelSpec("polyharmonic", 1.0, 2), "ph_k1_e001": O.KernelSpec("polyharmonic", 0.01, 1),
    "mq_e07": O.KernelSpec("multiquadric", 0.7, 0),
}


@pytest.mark.parametrize("name", sorted(KERNEL_CASES))
def test_kernel_values_match_reference(name):
    g = load("kernels.npz")
    x, s = T64(g["x"]), T64(g["s"])
    k = KERNEL_CASES[name]
    close(k(x, s).numpy(), g[name])
    if name + "_single" in g.files:
        xs = x[0] if k.kind != "multiquadric" else x.reshape(5, -1)[0]
        sp = s if k.kind != "multiquadric" else s.reshape(7, -1)
        close(k(xs, sp).numpy(), g[name + "_single"])
    # the closed forms the CUDA kernels implement agree too
    rho = ((x.reshape(5, 1, -1) - s.reshape(1, 7, -1)) ** 2).sum(-1).numpy()
    kv, _ = A.radial(k.kind, k.a, k.n, rho)
    close(kv, g[name], 1e-9)  # cdist's |x|^2+|s|^2-2xs expansion is not used at these sizes, but allow slack


@pytest.mark.parametrize("name", ["rq_g10_p2", "ph_k1_e1", "ph_k3_e05", "mq_e07"])
def test_kernel_gradients_match_reference(name):
    g = load("kernels.npz")
    x, s, w = T64(g["x"]), T64(g["s"]), T64(g["w"])
    k = KERNEL_CASES[name]
    xv = (x if name != "mq_e07" else x.reshape(5, -1)).clone().requires_grad_(True)
    (k(xv, s) @ w).sum().backward()
    close(xv.grad.numpy(), g[name + "_gradx"])
    _, gx = A.score_grad_features(x.reshape(5, -1).numpy(), s.reshape(7, -1).numpy(), w.numpy(), k.kind, k.a, k.n)
    close(gx, np.asarray(g[name + "_gradx"]).reshape(5, -1), 1e-9)


FK_TOL = {"se2": 2e-6, "se3": 2e-6, "baxter_dual": 2e-6}  # float32-only maps in the reference


@pytest.mark.parametrize("name", ["planar2", "planar3", "planar7", "se2", "se3", "baxter", "baxter_right", "baxter_dual",
                                  "panda", "panda5", "dual_panda"])
def test_fk_matches_reference(name):
    g = load("fk.npz")
    robot = P.make_robot("baxter" if name == "baxter_right" else name)
    fk = P.oracle_fk(robot)
    q = T64(g[name + "_q"]).requires_grad_(True)
    x = fk(q)
    tol = FK_TOL.get(name, 1e-12)
    close(x.detach().numpy(), g[name + "_x"], tol)
    (x * T64(g[name + "_gx"])).sum().backward()
    close(q.grad.numpy(), g[name + "_gq"], max(tol, 1e-11))


def _analytic_fk(robot, q):
    from diffco_b200 import _lib

    d = robot.fk_desc
    kp = lambda rows: np.array([[d.keypoints[r][j] for j in range(d.n_keypoints)] for r in range(rows)])
    L = np.array([d.link_length[i] for i in range(d.n_links)])
    if d.type == _lib.DC_FK_PLANAR_CHAIN:
        return A.planar_chain(q, L)
    if d.type == _lib.DC_FK_SE2_BODY:
        return A.se2_body(q, kp(2))
    if d.type == _lib.DC_FK_SE3_BODY:
        return A.se3_body(q, kp(3))
    if d.type == _lib.DC_FK_SE2_BASE_PLANAR_ARM:
        return A.se2_base_planar_arm(q, kp(2), L)
    arms = []
    for a in range(d.n_arms):
        arm = d.arms[a]
        J = arm.n_joints
        base = np.eye(4)
        base[:3] = np.array([arm.base[i] for i in range(12)]).reshape(3, 4)
        arms.append(dict(a=[arm.a[i] for i in range(J)], d=[arm.d[i] for i in range(J)], s_alpha=[arm.s_alpha[i] for i in range(J)],
                         c_alpha=[arm.c_alpha[i] for i in range(J)], theta0=[arm.theta0[i] for i in range(J)],
                         mask=[arm.out_slot[i] >= 0 for i in range(J)], joint_index=[arm.joint_index[i] for i in range(J)],
                         base=base, offset=[arm.offset[i] for i in range(3)],
                         tool_points=[[arm.tool[t][r] for r in range(3)] for t in range(arm.n_tool)] or None))
    interleave = d.n_arms == 2 and min(s for s in d.arms[1].out_slot[: d.arms[1].n_joints] if s >= 0) == 1
    return A.dh_multi(q, arms, interleave)


@pytest.mark.parametrize("name", P.ROBOTS)
def test_analytic_jacobian_products_match_autograd(name):
    """The closed-form J^T products (what the CUDA epilogue computes) equal autograd through the oracle FK."""
    gen = torch.Generator().manual_seed(5)
    robot = P.make_robot(name)
    q = P.sample_configs(robot, 5, gen)
    qv = q.clone().requires_grad_(True)
    x = P.oracle_fk(robot)(qv)
    gx = torch.randn(x.shape, generator=gen, dtype=torch.float64)
    (x * gx).sum().backward()
    pts, vjp = _analytic_fk(robot, q.numpy())
    close(pts, x.detach().numpy(), 1e-12)
    # BaxterDualArmFK's float32-rounded base rotations are orthonormal only to ~3e-8, and the cross-product form of
    # the revolute-joint Jacobian assumes an exact rotation: 2e-8 relative, far inside the 1e-5 parity gate.
    close(vjp(gx.numpy()), qv.grad.numpy(), 1e-7 if name == "baxter_dual" else 1e-11)


def _oracle_perceptron(g, tag, dof):
    robot = P.make_robot(f"planar{dof}")
    fk = P.oracle_fk(robot)
    X, y = T64(g[f"{tag}_X"]), T64(g[f"{tag}_y"])
    kern = O.KernelSpec("rq", 10.0, 2)
    return robot, fk, kern, O.train_perceptron(X, y, kern, transform=fk, beta=1.0, max_iteration=len(X))


@pytest.mark.parametrize("tag,dof", [("p2", 2), ("p7", 7)])
def test_training_selects_reference_supports(tag, dof):
    g = load("perceptron.npz")
    _, fk, kern, perc = _oracle_perceptron(g, tag, dof)
    assert perc.support_index.tolist() == g[f"{tag}_idx"].tolist()  # bit-exact index selection
    close(perc.gains.numpy(), g[f"{tag}_gains"], 1e-9)
    close(perc.hypothesis.numpy(), g[f"{tag}_hyp"], 1e-9)
    close(perc.kernel_matrix.numpy(), g[f"{tag}_K"], 1e-12)
    nodes = O.fit_poly(perc, O.KernelSpec("polyharmonic", 1.0, 1), target="label")
    close(nodes.numpy(), g[f"{tag}_nodes"], 1e-7)


@pytest.mark.parametrize("tag,dof", [("p2", 2), ("p7", 7)])
def test_scores_and_gradients_match_reference(tag, dof):
    g = load("perceptron.npz")
    robot = P.make_robot(f"planar{dof}")
    fk = P.oracle_fk(robot)
    X = T64(g[f"{tag}_X"])
    S = X[torch.from_numpy(g[f"{tag}_idx"])]
    St = fk(S)
    gains, nodes, Q = T64(g[f"{tag}_gains"]), T64(g[f"{tag}_nodes"]), T64(g[f"{tag}_Q"])
    rq, ph = O.KernelSpec("rq", 10.0, 2), O.KernelSpec("polyharmonic", 1.0, 1)
    s, gs = O.score_and_grad(lambda q: O.score_original(q, fk, rq, St, gains), Q)
    close(s.numpy(), g[f"{tag}_score"])
    close(gs.numpy(), g[f"{tag}_score_grad"], 1e-11)
    p, gp = O.score_and_grad(lambda q: O.poly_score(q, fk, ph, St, nodes), Q)
    close(p.numpy(), g[f"{tag}_poly"], 1e-11)
    close(gp.numpy(), g[f"{tag}_poly_grad"], 1e-10)
    assert O.score_original(Q[5], fk, rq, St, gains).shape == g[f"{tag}_single_score"].shape == ()
    assert O.poly_score(Q[5], fk, ph, St, nodes).shape == g[f"{tag}_single_poly"].shape == (1, 1)
    # closed form == autograd, including the r == 0 query (Q[0] coincides with a support)
    pts, vjp = _analytic_fk(robot, Q.numpy())
    sc, gx = A.score_grad_features(pts.reshape(len(Q), -1), St.reshape(len(St), -1).numpy(), nodes.numpy(), "polyharmonic", 1.0, 1)
    close(sc, g[f"{tag}_poly"], 1e-10)
    close(vjp(gx), g[f"{tag}_poly_grad"], 1e-9)


def test_jump_start_matches_reference():
    g = load("perceptron.npz")
    robot = P.make_robot("planar7")
    fk = P.oracle_fk(robot)
    kern = O.KernelSpec("rq", 10.0, 2)
    _, _, _, perc = _oracle_perceptron(g, "p7", 7)
    Xu, yu, exist = T64(g["p7u_X"]), T64(g["p7u_y"]), torch.from_numpy(g["p7u_exist"])
    gains0, h0, K0 = O.jump_start(perc, Xu, yu, exist, kern, transform=fk)
    close(gains0.numpy(), g["p7u_gains0"], 1e-9)
    close(h0.numpy(), g["p7u_h0"], 1e-9)
    close(K0.numpy(), g["p7u_K0"], 1e-12)
    upd = O.train_perceptron(Xu, yu, kern, transform=fk, beta=1.0, max_iteration=len(Xu), init=(gains0, h0, K0))
    assert upd.support_index.tolist() == g["p7u_idx"].tolist()
    close(upd.gains.numpy(), g["p7u_gains"], 1e-8)


def test_multiclass_matches_reference():
    g = load("multiclass.npz")
    robot = P.make_robot("baxter")
    fk = P.oracle_fk(robot)
    X, Y = T64(g["X"]), T64(g["Y"])
    rq = O.KernelSpec("rq", 10.0, 2)
    kcfg = lambda xi, Xall: rq(fk(xi[None, :]).reshape(1, -1), fk(Xall).reshape(len(Xall), -1))
    perc = O.train_multi_perceptron(X, Y, kcfg, beta=1.0, max_iteration=len(X))
    assert perc.support_index.tolist() == g["idx"].tolist()
    close(perc.gains.numpy(), g["gains"], 1e-9)
    close(perc.hypothesis.numpy(), g["hyp"], 1e-9)
    mq = O.KernelSpec("multiquadric", 1.0, 0)
    nodes = O.fit_poly_multi(perc, mq, fkine=fk, target="label")
    close(nodes.numpy(), g["nodes"], 1e-7)
    Q, go = T64(g["Q"]), T64(g["go"])
    S = fk(perc.support_points).reshape(len(perc.support_points), -1)
    s, gs = O.score_and_grad(lambda q: O.score_original(q, lambda z: fk(z).reshape(len(z), -1), rq, S, perc.gains), Q, go)
    close(s.numpy(), g["score"], 1e-11)
    close(gs.numpy(), g["score_grad"], 1e-10)
    r, gr = O.score_and_grad(lambda q: O.multi_rbf_score(q, fk, mq, S, T64(g["nodes"])), Q, go)
    close(r.numpy(), g["rbf"], 1e-10)
    close(gr.numpy(), g["rbf_grad"], 1e-9)


def test_optimizer_replay_queries_match_reference():
    g = load("optim_replay.npz")
    robot = P.make_robot("planar7")
    fk = P.oracle_fk(robot)
    St = fk(T64(g["support_points"]))
    nodes = T64(g["nodes"])
    ph = O.KernelSpec("polyharmonic", 1.0, 1)
    margin = float(g["safety_margin"])
    for j in range(int(g["n_kept"])):
        p = T64(g[f"call{j}_p"]).requires_grad_(True)
        sc = O.poly_score(p, fk, ph, St, nodes)
        close(sc.detach().numpy(), g[f"call{j}_score"], 1e-10)
        torch.clamp(sc - margin, min=0).sum().backward()
        close(p.grad.numpy(), g[f"call{j}_grad"], 1e-9)
    dense = O.dense_path(T64(g["con_p"]), float(g["max_speed"]))
    assert dense.shape[1] == 7 and torch.equal(dense[0], T64(g["con_p"])[0])


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("dense", [False, True])
def test_weighted_step_loop_matches_reference(dense):
    """oracle.weighted_step (the restated optim.py:706-752 loop) against the path the reference's Weighted.step produced
    on the same Baxter model (tests/golden/weighted.npz): same number of iterations, same waypoints."""
    import math

    g = load("weighted.npz")
    robot = P.make_robot("baxter")
    fk = P.oracle_fk(robot)
    St = fk(T64(g["support_points"]))
    nodes = T64(g["nodes"])
    ph = O.KernelSpec("polyharmonic", 1.0, 1)
    cw, mmw, jlw = (float(v) for v in g["weights"])
    x, steps = O.weighted_step(
        T64(g["init"]), fk, lambda q: O.poly_score(q, fk, ph, St, nodes), robot.limits.double(),
        lambda q: (math.pi + q) % (2 * math.pi) - math.pi, maxiter=int(g["maxiter"]), collision_weight=cw,
        max_move_weight=mmw, joint_limit_weight=jlw, safety_bias=float(g["safety_bias"]), max_speed=float(g["max_speed"]),
        lr=float(g["lr"]), dense_check=dense, mask=torch.from_numpy(g["mask"]))
    assert steps == int(g[f"steps_dense{int(dense)}"])
    close(x, g[f"x_dense{int(dense)}"], 1e-9)


def test_oracle_against_live_reference():
    """Fresh random problem, larger than the fixtures, straight against the imported reference."""
    import warnings

    warnings.filterwarnings("ignore")
    ns = ref_loader.load()
    gen = torch.Generator().manual_seed(77)
    robot = P.make_robot("panda")
    ref_robot = ns.model.PandaFK()
    q = P.sample_configs(robot, 64, gen)
    S = P.sample_configs(robot, 300, gen)
    w = torch.randn(300, generator=gen, dtype=torch.float64)
    fk = P.oracle_fk(robot)
    close(fk(q).numpy(), ref_robot.fkine(q).numpy(), 1e-12)
    dc = ns.kernel_perceptrons.DiffCo(kernel_func=ns.kernel.RQKernel(10.0), transform=ref_robot.fkine)
    dc.support_transformed, dc.gains = ref_robot.fkine(S), w
    qv = q.clone().requires_grad_(True)
    s_ref = dc.score(qv)
    s_ref.sum().backward()
    s, gq = O.score_and_grad(lambda z: O.score_original(z, fk, O.KernelSpec("rq", 10.0, 2), fk(S), w), q)
    close(s.numpy(), s_ref.detach().numpy(), 1e-12)
    close(gq.numpy(), qv.grad.numpy(), 1e-11)


def test_checker_fit_update_verify_match_reference():
    """oracle.checker_fit / checker_rates against the UNMODIFIED reference's ForwardKinematicsDiffCo (checkers.npz)."""
    g = load("checkers.npz")
    robot = P.make_robot("planar7")
    fk = P.oracle_fk(robot)
    kern = O.KernelSpec("rq", 10.0, 2)
    X = T64(g["X"])
    labels = (P.circle_labels(robot, X) > 0).double()
    torch.manual_seed(5)
    perc, bias, q_verify, y_verify = O.checker_fit(X, labels, kern, fk, verify_ratio=0.2)
    close(perc.support_points.numpy(), g["fit_support_points"], 1e-15)
    close(perc.rbf_nodes.numpy(), g["fit_nodes"], 1e-6)  # solve() of an ill-conditioned Polyharmonic Gram matrix: the
    # reference lays the features out (N, 3, L), a different summation order inside cdist
    close(q_verify.numpy(), g["fit_q_verify"], 1e-15)
    close(np.array(float(bias)), g["fit_bias"], 1e-7)
    rates = O.checker_rates(perc, fk, q_verify, y_verify, bias)
    close(np.array([float(v) for v in rates]), g["fit_rates"], 1e-6)
    Q = T64(g["Q"])
    ph = O.KernelSpec("polyharmonic", 1.0, 1)
    close((O.poly_score(Q, fk, ph, perc.support_transformed, perc.rbf_nodes) + bias).numpy(), g["fit_collision_score"].reshape(64, 1), 1e-7)
    Xu, exist = T64(g["upd_X"]), torch.from_numpy(g["upd_exist"])
    lab_u = (P.circle_labels(robot, Xu) > 0).double()
    perc2, bias2, _, _ = O.checker_fit(Xu, lab_u, kern, fk, verify_ratio=0, q_verify_fallback=T64(g["upd_q_verify"]),
                                       init_from=perc, exist_mask=exist)
    close(perc2.support_points.numpy(), g["upd_support_points"], 1e-15)
    close(np.array(float(bias2)), g["upd_bias"], 1e-7)
    close((O.poly_score(Q, fk, ph, perc2.support_transformed, perc2.rbf_nodes) + bias2).numpy(), g["upd_collision_score"], 1e-7)
    lab_q = 2 * (P.circle_labels(robot, Q) > 0).double() - 1
    close(np.array([float(v) for v in O.checker_rates(perc2, fk, Q, lab_q, bias2)]), g["upd_verify_rates"], 1e-6)


def test_line_and_temporal_kernels_match_reference():
    """kernel.py:145-202 (TemporalFKKernel, LineKernel, LineFKKernel) — restatement vs the reference's matrices and the
    autograd gradient of sum_n w_n k(x, s_n) w.r.t. the raw rows."""
    g = load("line_temporal.npz")
    fk = lambda q: O.fk_planar_chain(q, torch.ones(7, dtype=torch.float64))
    w = T64(g["w"])
    x, s = T64(g["t_x"]), T64(g["t_s"])
    for name in ("t_a", "t_b"):
        gx, px, gt, pt, al = g[name + "_params"]
        xv = x.clone().requires_grad_(True)
        km = O.temporal_fk_kernel(xv, s, fk, gx, int(px), gt, int(pt), al)
        (km @ w).sum().backward()
        close(km.detach(), g[name + "_K"])
        close(xv.grad, g[name + "_grad"], 1e-11)
        close(O.temporal_fk_kernel(x[0], s, fk, gx, int(px), gt, int(pt), al), g[name + "_K_single"])
    xl, sl = T64(g["l_x"]), T64(g["l_s"])
    xv = xl.clone().requires_grad_(True)
    km = O.line_fk_kernel(xv, sl, fk, 10.0)
    (km @ w).sum().backward()
    close(km.detach(), g["l_K"])
    close(xv.grad, g["l_grad"], 1e-11)
    close(O.line_fk_kernel(xl[0], sl, fk, 10.0), g["l_K_single"])
    close(O.line_kernel(xl, sl, lambda a, b: O.rq_kernel(a, b, 2.0)), g["lk_K"])


@pytest.mark.parametrize("name", ["arm7", "torso", "panda"])
def test_urdf_tree_fk_matches_reference(name):
    """rigid_body.py:86-141 through urdf_interface.py:517-553 (float32 in the reference): the float64 restatement run on the
    joint program this package compiled — for arm7 / torso compiled HERE from tests/data/*.urdf (so the XML parser and the
    compiler are pinned too; the golden's tree was built by an independent parse in oracle/make_golden.py), for the Panda
    from the descriptor bytes stored in the golden (the reference's file does not travel)."""
    from diffco_b200 import _lib
    from diffco_b200.collision_interfaces import URDFRobot

    g = load("urdf.npz")
    stored = _lib.FkDesc.from_buffer_copy(g[name + "_desc"].tobytes())
    if name != "panda":
        base = None if name == "arm7" else torch.tensor([[0.0, -1.0, 0.0, 0.3], [1.0, 0.0, 0.0, -0.2], [0.0, 0.0, 1.0, 0.1],
                                                          [0, 0, 0, 1.0]])
        path = os.path.join(os.path.dirname(__file__), "data", {"arm7": "arm7_gripper.urdf", "torso": "torso_two_arms.urdf"}[name])
        robot = URDFRobot(path, base_transform=base)
        assert bytes(robot.fk_desc) == bytes(stored)
        assert robot.unique_position_link_names == list(g[name + "_unique"]) and robot.node_names == list(g[name + "_nodes"])
        close(robot.joint_limits, g[name + "_limits"], 1e-7)
        desc = robot.fk_desc
    else:
        desc = stored
    nodes = P.tree_nodes_from_desc(desc)
    q = T64(g[name + "_q"]).requires_grad_(True)
    x, frames = O.fk_joint_tree(q, nodes, desc.n_points)
    close(x.detach().transpose(1, 2), g[name + "_x"], 3e-6)  # reference layout (B, 3, L), float32 arithmetic
    (x.transpose(1, 2) * T64(g[name + "_gx"])).sum().backward()
    close(q.grad, g[name + "_gq"], 1e-5)
    order = [list(g[name + "_nodes"]).index(k) for k in g[name + "_links"]]
    close(torch.stack([frames[i][1] for i in order], 1).detach(), g[name + "_trans"], 3e-6)
    close(torch.stack([frames[i][0] for i in order], 1).detach(), g[name + "_rot"], 3e-6)
