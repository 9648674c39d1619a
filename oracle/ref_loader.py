"""Import the *unmodified* reference (ucsdarclab/diffco) from /root/reference for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffco_b200/`` may import this module.  It is used by
``oracle/make_golden.py`` (fixture generation, in the build container) and by the ``-m "not gpu"``
tests that pin ``oracle/diffco_oracle.py`` against the reference when ``/root/reference`` exists.
``/root/reference`` does not exist on the GPU box, so everything here degrades to "unavailable".

Recipe: the reference's ``diffco/__init__.py`` pulls in fcl/trimesh/yourdfpy/matplotlib,
none of which are installed.  We therefore register an empty package object for ``diffco`` whose
``__path__`` points at the reference sources (so ``__init__`` is never executed), register empty stub
modules for the plotting / geometry dependencies that are only touched off the hot path, and import
the hot-path modules individually.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIFFCO_REFERENCE_ROOT", "/root/reference")
_STUBS = ("matplotlib", "matplotlib.pyplot", "fcl", "trimesh", "seaborn")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "diffco", "kernel.py"))


def _stub(name):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__["__stub__"] = True
    sys.modules[name] = m
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, m)


def load():
    """Return a namespace with the reference hot-path modules (kernel, utils, model, ...)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for s in _STUBS:
        _stub(s)
    if "diffco" not in sys.modules or not getattr(sys.modules["diffco"], "__refpkg__", False):
        pkg = types.ModuleType("diffco")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "diffco")]
        pkg.__refpkg__ = True
        sys.modules["diffco"] = pkg
    ns = types.SimpleNamespace()
    for sub in ("kernel", "utils", "kernel_perceptrons", "model", "robot_fkine", "optim"):
        setattr(ns, sub, importlib.import_module("diffco." + sub))
    return ns


def load_legacy():
    """Legacy multi-class perceptron (diffco/deprecated/MultiDiffCo.py) with the FKKernel shim.

    The live ``kernel.FKKernel`` raises in its constructor (kernel.py:131-133); the legacy scripts were
    written against its pre-deprecation behaviour (kernel.py:137-143), which the shim reproduces.
    """
    ns = load()
    if "diffco_legacy" not in sys.modules:
        pkg = types.ModuleType("diffco_legacy")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "diffco", "deprecated")]
        sys.modules["diffco_legacy"] = pkg
        sys.modules["diffco_legacy.kernel"] = ns.kernel
        pkg.kernel = ns.kernel
    ns.legacy_DiffCo = importlib.import_module("diffco_legacy.DiffCo")
    ns.legacy_MultiDiffCo = importlib.import_module("diffco_legacy.MultiDiffCo")

    class FKKernelShim(ns.kernel.KernelFunc):
        def __init__(self, fkine, rq_kernel):
            self.fkine = fkine
            self.rq_kernel = rq_kernel

        def __call__(self, xs, x_primes=None, x_primes_controls=None):
            if xs.ndim == 1:
                xs = xs[None, :]
            xc = self.fkine(xs).reshape(len(xs), -1)
            if x_primes_controls is None:
                x_primes_controls = self.fkine(x_primes).reshape(len(x_primes), -1)
            return self.rq_kernel(xc, x_primes_controls)

    ns.FKKernelShim = FKKernelShim
    return ns
