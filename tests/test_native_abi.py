"""CPU tests of the drop-in boundary: libdiffco_b200.so loads without a GPU, exports every symbol declared in
include/diffco_b200.h, the ctypes mirror of the descriptor structs has the C layout, and argument validation that
needs no device answers with status codes.  No compute calls here (those are the -m gpu tests)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "diffco_b200.h")


@pytest.fixture(scope="module")
def lib():
    from diffco_b200 import _lib, build

    build.build()  # no-op when the in-tree .so is current; nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|void|int64_t|double|const char\*)\s+(dc_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("dc_score_grad", "dc_kernel_matrix", "dc_fk_forward", "dc_fk_vjp", "dc_perceptron_train", "dc_pack_supports",
                 "dc_score_grad_host", "dc_host_pipeline_create", "dc_host_pipeline_destroy"):
        assert must in names, names


def test_library_exports_every_declared_symbol(lib):
    from diffco_b200 import _lib

    names = declared_functions()
    assert sorted(_lib.PROTOTYPES) == names  # the binding lists exactly what the header declares
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.dc_abi_version() == int(re.search(r"#define DC_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dc_\w+)", out))
    assert set(names) <= exported


def test_struct_layouts_match_the_c_header(tmp_path):
    from diffco_b200 import _lib

    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "diffco_b200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(dc_dh_arm), sizeof(dc_fk_desc),"
        " sizeof(dc_kernel_desc), sizeof(dc_supports), offsetof(dc_fk_desc, link_length), offsetof(dc_fk_desc, keypoints),"
        " offsetof(dc_fk_desc, arms), offsetof(dc_dh_arm, base), offsetof(dc_supports, tc_blob), sizeof(dc_traj_params),"
        " offsetof(dc_traj_params, limits), offsetof(dc_traj_params, wrap)); return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.DhArm), C.sizeof(_lib.FkDesc), C.sizeof(_lib.KernelDesc), C.sizeof(_lib.Supports),
            _lib.FkDesc.link_length.offset, _lib.FkDesc.keypoints.offset, _lib.FkDesc.arms.offset, _lib.DhArm.base.offset,
            _lib.Supports.tc_blob.offset, C.sizeof(_lib.TrajParams), _lib.TrajParams.limits.offset, _lib.TrajParams.wrap.offset]
    assert got == want


def test_layout_helper_and_validation_without_a_device(lib):
    f_pad, row = C.c_int32(), C.c_int32()
    assert lib.dc_supports_layout(14, 1, 0, C.byref(f_pad), C.byref(row)) == 0
    assert (f_pad.value, row.value) == (14, 16)
    assert lib.dc_supports_layout(12, 4, 0, C.byref(f_pad), C.byref(row)) == 0
    assert (f_pad.value, row.value) == (12, 16)
    assert lib.dc_supports_layout(21, 1, 1, C.byref(f_pad), C.byref(row)) == 0
    assert (f_pad.value, row.value) == (22, 24)
    assert lib.dc_supports_layout(0, 1, 0, None, None) == -1
    assert lib.dc_supports_layout(65, 1, 0, None, None) == -1
    assert lib.dc_supports_layout(4, 9, 0, None, None) == -1
    assert lib.dc_supports_layout(4, 1, 7, None, None) == -1
    assert lib.dc_status_string(0) == b"ok" and lib.dc_status_string(-2) == b"unsupported configuration"
    assert lib.dc_score_grad(None, None, None, None, 1, None, 0, None, 0, None, 0, None) == -1
    assert lib.dc_launch_count() >= 0


def test_tensor_core_peer_and_trajectory_entry_points_validate_arguments(lib):
    """Argument checks of the entry points added with ABI version 2 — all of them return before touching a device."""
    from diffco_b200 import _lib

    n = C.c_int64()
    assert lib.dc_supports_tc_bytes(2000, 14, 1, _lib.DC_F32, C.byref(n)) == 0
    # 21 chunks of 96 supports: GEMM1 image, GEMM2 image + fp32 weights + chunk maximum, trailer
    assert n.value == 21 * (9216 + 12800) + 64
    assert lib.dc_supports_tc_bytes(2000, 15, 1, _lib.DC_F32, C.byref(n)) == 0    # 15 <= F <= 30: operand groups of 32 K slots
    assert n.value == 21 * (96 * 96 * 2 + 12 * 2048 + 512) + 64
    assert lib.dc_supports_tc_bytes(2000, 31, 1, _lib.DC_F32, C.byref(n)) == -2   # F > 30
    assert lib.dc_supports_tc_bytes(2000, 14, 4, _lib.DC_F32, C.byref(n)) == -2   # multi-class
    assert lib.dc_supports_tc_bytes(2000, 14, 1, _lib.DC_F64, C.byref(n)) == -2   # float64
    assert lib.dc_supports_tc_bytes(0, 14, 1, _lib.DC_F32, C.byref(n)) == -1
    rq = _lib.KernelDesc(_lib.DC_K_RQ, 2, 10.0)
    assert lib.dc_pack_supports_tc(None, None, 10, 14, C.byref(rq), None, None) == -1
    assert lib.dc_pack_supports_tc(64, 64, 10, 14, C.byref(rq), 72, None) == -1   # blob not 128-byte aligned
    assert lib.dc_pack_supports_tc(64, 64, 10, 14, None, 128, None) == -1         # the image is built for one kernel
    ph = _lib.KernelDesc(_lib.DC_K_POLYHARMONIC, 1, 1.0)
    assert lib.dc_pack_supports_tc(64, 64, 10, 14, C.byref(ph), 128, None) == -2  # RQKernel(p = 2) only
    assert lib.dc_supports_tc_info(None, 10, 14, None, None) == -1
    saved = [lib.dc_get_option(k) for k in (1, 2, 3, 4)]
    try:
        assert lib.dc_set_option(_lib.DC_OPT_TC_ERR_COEF, 1e-6) == 0 and lib.dc_get_option(_lib.DC_OPT_TC_ERR_COEF) == 1e-6
        assert lib.dc_set_option(_lib.DC_OPT_TC_ENABLE, 0.0) == 0 and lib.dc_get_option(_lib.DC_OPT_TC_ENABLE) == 0.0
        assert lib.dc_set_option(_lib.DC_OPT_TC_TOL_PAIR, 0.0) == -1 and lib.dc_set_option(_lib.DC_OPT_TC_MIN_BATCH, 0.5) == -1
        assert lib.dc_set_option(77, 1.0) == -1
    finally:
        for k, v in zip((1, 2, 3, 4), saved):
            assert lib.dc_set_option(k, v) == 0
    assert lib.dc_last_score_kernel() in (-1, 0, 1, 2)
    tab = _lib.PeerTable()
    assert lib.dc_peer_barrier(C.byref(tab), 0, 2, 1, None) == -1       # unmapped flag arrays
    assert lib.dc_peer_barrier(C.byref(tab), 3, 2, 1, None) == -1 and lib.dc_peer_barrier(None, 0, 1, 1, None) == -1
    assert lib.dc_peer_alloc(0, None, None) == -1 and lib.dc_peer_open(None, None) == -1
    assert lib.dc_peer_close(None) == 0 and lib.dc_peer_free(None) == 0
    fk, kd, sv = _lib.FkDesc(), _lib.KernelDesc(_lib.DC_K_RQ, 2, 10.0), _lib.Supports()
    assert lib.dc_score_grad_bcast(C.byref(fk), C.byref(kd), C.byref(sv), 64, 8, C.byref(tab), 0, 0, 1, None, None) == -1
    assert lib.dc_score_grad_bcast(C.byref(fk), C.byref(kd), C.byref(sv), 64, 8, C.byref(tab), 2, 0, 1, None, None) == -1  # null outs
    prm = _lib.TrajParams()
    assert lib.dc_traj_step(None, C.byref(prm), 8, 0, None, None, None, None, None, None, None, None, None) == -1
    fk.type, fk.dof, fk.n_points, fk.point_dim, fk.n_links = _lib.DC_FK_PLANAR_CHAIN, 3, 3, 2, 3
    assert lib.dc_traj_step(C.byref(fk), C.byref(prm), 1, 0, 64, None, None, None, 64, 64, 64, 64, None) == -1   # < 2 waypoints
    assert lib.dc_traj_step(C.byref(fk), C.byref(prm), 8, 0, 64, 64, None, None, 64, 64, 64, 64, None) == -1    # score without grad
    assert lib.dc_traj_step(C.byref(fk), C.byref(prm), 8, 0, 64, None, None, None, 64, 64, 64, 64, None) == -1  # lr == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: scoring without a CUDA device raises instead of silently computing elsewhere."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=3)
    dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine)
    dc.support_points = torch.zeros(4, 3)
    dc.support_transformed = torch.zeros(4, 3, 2)
    dc.gains = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        dc.score(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        robot.fkine(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        dc.train(torch.zeros(4, 3), torch.ones(4))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under diffco_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "diffco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f
    code = "import sys; import diffco_b200, diffco_b200.functional; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_autograd_plumbing_with_a_stub_evaluator():
    """Host logic of functional.differentiable_score: plain backward, batched backward (jacobian(vectorize=True), the
    call optim.py:211-216 makes), no_grad, and a loud failure for second derivatives — with a stub evaluator standing
    in for the CUDA launch."""
    from diffco_b200 import functional as Fn

    A = torch.randn(3, 5, dtype=torch.float64)

    def evaluator(q, want_jac):  # score[b,c] = sum_d A[c,d] q[b,d]^2 ; jac[b,c,d] = 2 A[c,d] q[b,d]
        s = (q.detach() ** 2) @ A.T
        return s, (2 * A[None] * q.detach()[:, None, :] if want_jac else None)

    q = torch.randn(7, 5, dtype=torch.float64, requires_grad=True)
    go = torch.randn(7, 3, dtype=torch.float64)
    s = Fn.differentiable_score(q, evaluator)
    (s * go).sum().backward()
    ref = torch.autograd.grad((((q**2) @ A.T) * go).sum(), q)[0]
    assert torch.allclose(q.grad, ref)
    f = lambda z: Fn.differentiable_score(z, evaluator).sum(1)
    jac = torch.autograd.functional.jacobian(f, q, vectorize=True, strategy="reverse-mode")
    jref = torch.autograd.functional.jacobian(lambda z: ((z**2) @ A.T).sum(1), q)
    assert torch.allclose(jac, jref)
    with torch.no_grad():
        assert not Fn.differentiable_score(q, evaluator).requires_grad
    (g1,) = torch.autograd.grad(Fn.differentiable_score(q, evaluator).sum(), q, create_graph=True)
    with pytest.raises(RuntimeError):
        g1.sum().backward()


def test_integration_md_stub_matches_the_binding():
    """The ctypes stub INTEGRATION.md shows a reference maintainer is executable and its structures have the sizes of the
    real binding (diffco_b200/_lib.py), which the layout test above pins to the header."""
    import ctypes
    import re

    from diffco_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    stub = re.findall(r"```python\n(.*?)```", text, re.S)[0]
    stub = stub.replace('C.CDLL("libdiffco_b200.so")', f'C.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    for name in ("DhArm", "TreeNode", "FkDesc", "KernelDesc", "Supports"):
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(getattr(_lib, name)), name
