"""diffco_b200 — B200-native (sm_100a CUDA) implementation of DiffCo's batched collision-score hot path.

Public surface mirrors the reference package (diffco/__init__.py:1-4) for the path in scope: ``kernel``, ``model``,
``utils``, ``optim``, ``DiffCo``, ``MultiDiffCo`` and the high-level checkers ``RBFDiffCo`` / ``ForwardKinematicsDiffCo`` with an
injected ground truth (``routines`` and the geometry back ends are out of scope).
"""
from . import kernel, model, optim, utils  # noqa: F401
from .collision_checkers import CollisionChecker, ForwardKinematicsDiffCo, RBFDiffCo  # noqa: F401
from .kernel_perceptrons import DiffCo, MultiDiffCo  # noqa: F401

__all__ = ["kernel", "model", "utils", "optim", "DiffCo", "MultiDiffCo", "CollisionChecker", "RBFDiffCo",
           "ForwardKinematicsDiffCo"]
