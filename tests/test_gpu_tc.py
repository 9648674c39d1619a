"""GPU tests (-m gpu) of the tensor-core score kernel (csrc/dc_score_tc.cuh: tcgen05.mma kind::f16 with split operands,
TMEM accumulators, bulk-TMA operand blobs, exact FP32 evaluation of near pairs), called through the C ABI.

Checked against the float64 oracle (gate 1e-5 of max|ref|), against the FP32-pipe kernel on the same
inputs, and for the properties that do not depend on size: position independence, the near-pair path (queries on and
next to support vectors), out-of-range queries, the fused [score | grad] record and the host-buffer pipeline.
"""
import pytest
import torch

from oracle import diffco_oracle as O
from tests import problems as P
from tests.test_gpu_parity import cuda_support_set, kernel_pair, oracle_score_grad, rel

pytestmark = pytest.mark.gpu

TC, TQ, LS = 2, 1, 0


@pytest.fixture(scope="module")
def dev(cuda_device):
    from diffco_b200 import _lib

    _lib.load()
    return cuda_device


@pytest.fixture()
def lib():
    from diffco_b200 import _lib

    L = _lib.load()
    saved = [L.dc_get_option(k) for k in (1, 2, 3, 4)]
    yield L
    for k, v in zip((1, 2, 3, 4), saved):
        assert L.dc_set_option(k, v) == 0


def run(lib, robot, kfun, sv, qd, mode, tc, go=None, out=None):
    from diffco_b200 import functional as Fn

    assert lib.dc_set_option(1, 1.0 if tc else 0.0) == 0
    s, g = Fn.score_grad(robot.fk_desc, kfun.desc, sv, qd, mode, go, out)
    return s, g, lib.dc_last_score_kernel()


@pytest.mark.parametrize("rname", ["planar7", "planar2", "planar3", "se2", "baxter", "se2arm"])
def test_tensor_core_kernel_matches_oracle_and_fp32_kernel(rname, dev, lib):
    from diffco_b200 import _lib

    robot, S, W = P.synthetic_model(rname, 1501, 1, seed=310)
    gen = torch.Generator().manual_seed(311)
    q = P.sample_configs(robot, 9001, gen)
    # the near-pair path: exact coincidences, tiny and moderate offsets from support vectors
    q[10] = S[3]
    q[4000:4200] = S[:200] + 0.05 * torch.randn(200, robot.dof, generator=gen, dtype=torch.float64)
    q[6000:6400] = S[200:600] + 0.2 * torch.randn(400, robot.dof, generator=gen, dtype=torch.float64)
    q = q.float().double()  # the oracle sees exactly the float32 configurations the kernels get
    kfun, kspec = kernel_pair("rq")
    s_ref, g_ref = oracle_score_grad(robot, kspec, S, W, q)
    sv = cuda_support_set(robot, S, W, torch.float32, dev)
    assert sv.tc_blob is not None and sv.desc.tc_s2max > 0
    qd = q.to(device=dev, dtype=torch.float32)
    s, g, which = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True)
    print(f"{rname}: F = {sv.n_features}, max|s|^2 = {sv.desc.tc_s2max:.1f} -> {_lib.KERNEL_NAMES[which]}", flush=True)
    if rname == "planar7":  # the other maps are narrower than the kernel: the dispatcher may keep them on the FP32 pipe
        assert which == TC, _lib.KERNEL_NAMES[which]
    s0, g0, which0 = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=False)
    assert which0 == TQ
    es, eg = rel(s, s_ref), rel(g, g_ref)
    es0, eg0 = rel(s0, s_ref), rel(g0, g_ref)
    print(f"{rname}: tensor-core score {es:.2e} grad {eg:.2e} | fp32 kernel score {es0:.2e} grad {eg0:.2e}", flush=True)
    # The tensor-core path is held to 1e-5 everywhere: its exact near-pair path works on (hi, lo) feature pairs from the
    # float64 forward kinematics.  The FP32-pipe kernel sees the float32 features only; with se2arm's base translations
    # of up to 10 m (ulp 9.5e-7) that alone is ~1.1e-5 of the gradient maximum next to a support vector — the stated
    # float32-feature limit of that kernel, not a property of this one.
    gate0 = 1e-5 if rname != "se2arm" else 2e-5
    assert es0 <= gate0 and eg0 <= gate0
    assert es <= 1e-5 and eg <= 1e-5
    assert rel(s, s0) <= gate0 and rel(g, g0) <= gate0
    gate = 1e-5
    # score-only instantiation and the upstream gradient folded into the launch
    s1, none, which1 = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_NONE, tc=True)
    assert which1 == which and none is None and rel(s1, s_ref) <= 1e-5
    go = torch.randn(len(q), 1, generator=gen, dtype=torch.float64)
    _, g_ref2 = oracle_score_grad(robot, kspec, S, W, q, go)
    _, g2, _ = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True, go=go.to(device=dev, dtype=torch.float32))
    assert rel(g2, g_ref2) <= gate


@pytest.mark.parametrize("gamma,forced", [(10.0, True), (300.0, False)])
@pytest.mark.parametrize("rname", ["panda", "baxter_dual", "planar15"])
def test_wide_feature_maps_on_the_tensor_core_kernel(rname, gamma, forced, dev, lib):
    """15 <= F <= 30 (Panda 21, dual Baxter 24, a 15-link planar arm 30): operand groups of 32 K slots, one CTA per SM.  With the
    reference's default width (gamma = 10) these metre-scale maps are wider than the kernel — a large share of the pairs
    is "near" and the dispatcher rightly keeps the FP32-pipe kernels; the tensor-core instantiation is then forced
    (max|s|^2 declared unknown) to hold ITS arithmetic to the same 1e-5.  With a narrow kernel it is chosen on its own."""
    from diffco_b200 import _lib
    from diffco_b200 import kernel as K
    from oracle import diffco_oracle as O

    robot, S, W = P.synthetic_model(rname, 1501, 1, seed=410)
    gen = torch.Generator().manual_seed(411)
    q = P.sample_configs(robot, 8192 + 77, gen)
    q[10] = S[3]
    q[4000:4200] = S[:200] + 0.05 * torch.randn(200, robot.dof, generator=gen, dtype=torch.float64)
    q[6000:6400] = S[200:600] + 0.2 * torch.randn(400, robot.dof, generator=gen, dtype=torch.float64)
    q = q.float().double()
    S, W = S.float().double(), W.float().double()
    kfun, kspec = K.RQKernel(gamma), O.KernelSpec("rq", gamma, 2)
    s_ref, g_ref = oracle_score_grad(robot, kspec, S, W, q)
    sv = cuda_support_set(robot, S, W, torch.float32, dev, kfun=kfun)
    assert sv.tc_blob is not None and sv.n_features > 14
    if forced:
        sv.desc.tc_s2max = 0.0
    qd = q.to(device=dev, dtype=torch.float32)
    assert lib.dc_set_option(_lib.DC_OPT_TC_STATS, 1.0) == 0
    s, g, which = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True)
    near = lib.dc_get_option(_lib.DC_OPT_TC_STATS) / (len(q) * len(S))
    assert lib.dc_set_option(_lib.DC_OPT_TC_STATS, 0.0) == 0
    assert which == TC, _lib.KERNEL_NAMES[which]
    s0, g0, which0 = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=False)
    assert which0 == (TQ if sv.n_features in (21, 24, 27) else LS)  # thread-per-query instantiations: F <= 16, 21, 24, 27
    es, eg, es0, eg0 = rel(s, s_ref), rel(g, g_ref), rel(s0, s_ref), rel(g0, g_ref)
    print(f"{rname} gamma {gamma}: F = {sv.n_features}, exact-path pairs {100 * near:.2f} % | tensor-core score {es:.2e} grad {eg:.2e} "
          f"| lane-split score {es0:.2e} grad {eg0:.2e}", flush=True)
    assert es <= 1e-5 and eg <= 1e-5
    s1, none, _ = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_NONE, tc=True)
    assert none is None and rel(s1, s_ref) <= 1e-5
    # position independence: the same rows in another order give the same bits
    perm = torch.randperm(len(q), generator=gen)
    sp, gp, _ = run(lib, robot, kfun, sv, qd[perm.to(dev)], _lib.DC_GRAD_SUM, tc=True)
    assert torch.equal(sp, s[perm.to(dev)]) and torch.equal(gp, g[perm.to(dev)])


@pytest.mark.parametrize("n_feat", [9, 14, 15, 22, 30])
def test_raw_feature_rows_on_the_tensor_core_kernel(n_feat, dev, lib):
    """transform=None (kernel_perceptrons.py:36; poly_score(transformed_point=...) :316-317): the rows ARE the features and
    gradients are w.r.t. them — both instantiations (F <= 14 and 15 <= F <= 30), records of up to 31 floats per query."""
    from diffco_b200 import _lib
    from diffco_b200 import functional as Fn
    from diffco_b200 import kernel as K

    gen = torch.Generator().manual_seed(500 + n_feat)
    S = (torch.randn(1203, n_feat, generator=gen, dtype=torch.float64) * 1.5).float().double()
    W = torch.randn(1203, 1, generator=gen, dtype=torch.float64).float().double()
    q = torch.randn(8192 + 31, n_feat, generator=gen, dtype=torch.float64) * 1.5
    q[100:300] = S[:200] + 0.02 * torch.randn(200, n_feat, generator=gen, dtype=torch.float64)
    q = q.float().double()
    kfun, kspec = K.RQKernel(6.0), O.KernelSpec("rq", 6.0, 2)
    f = lambda z: O.score_original(z, None, kspec, S, W)
    s_ref, g_ref = O.score_and_grad(f, q)
    sv = Fn.SupportSet(S.float().to(dev), W.float().to(dev), dev, kernel=kfun.desc)
    assert sv.tc_blob is not None
    sv.desc.tc_s2max = 0.0  # "unknown": skip the dispatcher's width test, this case is about the transform=None path
    fk = Fn.none_fk(n_feat)
    assert lib.dc_set_option(1, 1.0) == 0
    s, g = Fn.score_grad(fk, kfun.desc, sv, q.float().to(dev), _lib.DC_GRAD_SUM)
    assert lib.dc_last_score_kernel() == TC
    assert rel(s, s_ref.reshape(len(q), -1)) <= 1e-5 and rel(g, g_ref) <= 1e-5


def test_cfg2_full_size_sampled_rows_and_position_independence(dev, lib):
    """BASELINE.json configs[1] (7-DoF planar arm, 2000 SVs, batch 65536) on the tensor-core kernel: 512 sampled rows
    against the oracle, and bit-identical rows when the batch is permuted (a row's result may not depend on its tile)."""
    from diffco_b200 import _lib

    robot, S, W = P.synthetic_model("planar7", 2000, 1, seed=1234)
    gen = torch.Generator().manual_seed(1235)
    q = P.sample_configs(robot, 65536, gen)
    kfun, kspec = kernel_pair("rq")
    sv = cuda_support_set(robot, S, W, torch.float32, dev)
    qd = q.to(device=dev, dtype=torch.float32)
    s, g, which = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True)
    assert which == TC
    rows = torch.randperm(65536, generator=gen)[:512]
    s_ref, g_ref = oracle_score_grad(robot, kspec, S, W, q[rows])
    # the gate is relative to the largest value of the whole batch; sampled rows are compared on that scale
    smax, gmax = s.abs().max().item(), g.abs().max().item()
    assert (s[rows.to(dev)].double().cpu() - s_ref).abs().max().item() <= 1e-5 * smax
    assert (g[rows.to(dev)].double().cpu() - g_ref).abs().max().item() <= 1e-5 * gmax
    perm = torch.randperm(65536, generator=gen).to(dev)
    s2, g2, _ = run(lib, robot, kfun, sv, qd[perm].contiguous(), _lib.DC_GRAD_SUM, tc=True)
    assert torch.equal(s2, s[perm]) and torch.equal(g2, g[perm])
    # linearity in the weights: score(W) + score(2W) == score(3W) up to rounding
    sv3 = cuda_support_set(robot, S, 3.0 * W, torch.float32, dev)
    s3, _, which3 = run(lib, robot, kfun, sv3, qd, _lib.DC_GRAD_NONE, tc=True)
    assert which3 == TC and rel(s3, 3.0 * s) <= 2e-6


def test_fused_record_ragged_tiles_and_host_pipeline(dev, lib):
    from diffco_b200 import _lib
    from diffco_b200 import functional as Fn

    robot, S, W = P.synthetic_model("planar7", 777, 1, seed=320)
    gen = torch.Generator().manual_seed(321)
    kfun, kspec = kernel_pair("rq")
    sv = cuda_support_set(robot, S, W, torch.float32, dev)
    for B in (4096, 4097, 5000, 12345):
        q = P.sample_configs(robot, B, gen)
        s_ref, g_ref = oracle_score_grad(robot, kspec, S, W, q)
        qd = q.to(device=dev, dtype=torch.float32)
        out = torch.full((B, 1 + robot.dof), float("nan"), device=dev)
        s, g, which = run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True, out=out)
        assert which == TC
        assert rel(out[:, :1], s_ref) <= 1e-5 and rel(out[:, 1:], g_ref) <= 1e-5
        # strided record (row stride > C + D): the non-fused store path
        wide = torch.zeros((B, 12), device=dev)
        run(lib, robot, kfun, sv, qd, _lib.DC_GRAD_SUM, tc=True, out=wide[:, : 1 + robot.dof])
        assert torch.equal(wide[:, : 1 + robot.dof], out) and float(wide[:, 1 + robot.dof:].abs().max()) == 0.0
    # host buffers in, host buffers out (zero-copy over pinned memory)
    qh = q.float().pin_memory()
    oh = torch.empty((len(q), 1 + robot.dof), dtype=torch.float32).pin_memory()
    assert lib.dc_set_option(1, 1.0) == 0
    Fn.HostPipeline(dev).score_grad(robot.fk_desc, kfun.desc, sv, qh, oh, _lib.DC_GRAD_SUM)
    torch.cuda.synchronize()
    assert lib.dc_last_score_kernel() == TC
    assert rel(oh[:, :1], s_ref) <= 1e-5 and rel(oh[:, 1:], g_ref) <= 1e-5


def test_out_of_range_queries_and_large_weights(dev, lib):
    """Features far outside the supports' range cannot be represented in the f16 operand scaling: those rows take the
    exact FP32 path for every pair.  Weights are rescaled by a power of two, so their magnitude must not matter."""
    from diffco_b200 import _lib
    from diffco_b200 import functional as Fn

    gen = torch.Generator().manual_seed(330)
    S = torch.randn(600, 6, generator=gen, dtype=torch.float64)
    W = 1e6 * torch.randn(600, 1, generator=gen, dtype=torch.float64)
    q = torch.randn(4500, 6, generator=gen, dtype=torch.float64)
    q[7] *= 1e4
    q[4100] = torch.tensor([3e4, -2e4, 0, 0, 0, 0], dtype=torch.float64)
    q[11] = S[0]
    from diffco_b200 import kernel as K

    kfun, kspec = K.RQKernel(40.0), O.KernelSpec("rq", 40.0, 2)
    fk = Fn.none_fk(6)
    sv = Fn.SupportSet(S.float().to(dev), W.float().to(dev), dev, kernel=kfun.desc)
    f = lambda z: O.score_original(z, lambda t: t, kspec, S.float().double(), W.float().double())
    s_ref, g_ref = O.score_and_grad(f, q.float().double())
    assert lib.dc_set_option(1, 1.0) == 0
    s, g = Fn.score_grad(fk, kfun.desc, sv, q.float().to(dev), _lib.DC_GRAD_SUM)
    assert lib.dc_last_score_kernel() == TC
    assert rel(s, s_ref.reshape(-1, 1)) <= 1e-5 and rel(g, g_ref) <= 1e-5
    assert torch.isfinite(s).all() and torch.isfinite(g).all()


def test_dispatch_rules(dev, lib):
    from diffco_b200 import _lib
    from diffco_b200 import kernel as K

    robot, S, W = P.synthetic_model("planar7", 300, 1, seed=340)
    q = P.sample_configs(robot, 5000, torch.Generator().manual_seed(341)).to(device=dev, dtype=torch.float32)
    sv = cuda_support_set(robot, S, W, torch.float32, dev)
    rq = K.RQKernel(10.0)
    assert run(lib, robot, rq, sv, q, _lib.DC_GRAD_SUM, tc=True)[2] == TC
    assert run(lib, robot, rq, sv, q[:2000], _lib.DC_GRAD_SUM, tc=True)[2] != TC          # below DC_OPT_TC_MIN_BATCH
    assert run(lib, robot, K.RQKernel(0.05), sv, q, _lib.DC_GRAD_SUM, tc=True)[2] == TQ   # wide kernel: most pairs near
    assert run(lib, robot, K.RQKernel(10.0, 3), sv, q, _lib.DC_GRAD_SUM, tc=True)[2] != TC  # p != 2
    assert run(lib, robot, K.Polyharmonic(1, 1.0), sv, q, _lib.DC_GRAD_SUM, tc=True)[2] == TQ
    assert run(lib, robot, rq, sv, q, _lib.DC_GRAD_JAC, tc=True)[2] != TC or sv.n_class == 1
    assert lib.dc_set_option(99, 1.0) != 0 and lib.dc_set_option(2, -1.0) != 0
    # multi-class models have no tensor-core image; feature maps up to F = 30 do (Panda: F = 21)
    robot2, S2, W2 = P.synthetic_model("baxter", 200, 4, seed=342)
    assert cuda_support_set(robot2, S2, W2, torch.float32, dev).tc_blob is None
    robot3, S3, W3 = P.synthetic_model("panda", 200, 1, seed=343)
    assert cuda_support_set(robot3, S3, W3, torch.float32, dev).tc_blob is not None


@pytest.mark.parametrize("n_sv", [50, 96, 97, 193])
def test_support_counts_around_the_chunk_size(n_sv, dev, lib):
    """Half a chunk, exactly one chunk, one chunk + 1, two chunks + 1 (chunk = 96 supports; padding rows must vanish).
    (A single support is left to the dispatcher: with no pair inside the kernel's width the batch maximum is a far-field
    value, and the tensor-core path's bound — tol_pair of the WEIGHT per pair — is then ~1e-5 of that maximum.)"""
    from diffco_b200 import _lib

    robot, S, W = P.synthetic_model("planar7", n_sv, 1, seed=350 + n_sv)
    gen = torch.Generator().manual_seed(351)
    q = P.sample_configs(robot, 4200, gen)
    # a perceptron's supports ARE training configurations and queries lie among them: a third of the batch sits within the
    # kernel's width of a support, so the batch maxima are near-field values (with so few supports a purely uniform batch
    # is all far field: maxima ~1e-3, against which the tensor-core path's ABSOLUTE per-pair bound reads as ~1e-5)
    q[:1400] = S[torch.randint(n_sv, (1400,), generator=gen)] + 0.15 * torch.randn(1400, robot.dof, generator=gen, dtype=torch.float64)
    q = q.float().double()
    q[17] = S[0].float().double()
    kfun, kspec = kernel_pair("rq")
    s_ref, g_ref = oracle_score_grad(robot, kspec, S, W, q)
    sv = cuda_support_set(robot, S, W, torch.float32, dev)
    sv.desc.tc_s2max = 0.0  # "unknown": the dispatcher's width heuristic (few supports -> small max|s|^2) stays out of the way
    s, g, which = run(lib, robot, kfun, sv, q.to(device=dev, dtype=torch.float32), _lib.DC_GRAD_SUM, tc=True)
    assert which == TC
    assert rel(s, s_ref) <= 1e-5 and rel(g, g_ref) <= 1e-5
