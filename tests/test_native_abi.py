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
