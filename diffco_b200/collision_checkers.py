"""High-level checkers — host-side mirror of the reference's ``diffco/collision_checkers.py`` (``CollisionChecker``,
``RBFDiffCo``, ``ForwardKinematicsDiffCo``) on the device perceptron of this package.

Same methods, arguments and return values as the reference: ``fit`` (= ``perceptron.train`` +
``fit_poly(Polyharmonic(1, 1), 'label')``, collision_checkers.py:163-218), ``update`` (active learning: samples around
the support points + uniform samples, jump-start training, :220-252), ``verify`` (:254-290), ``collision`` /
``collision_score`` (= ``poly_score + safety_bias``, :292-303, :475-495), ``_calculate_safety_bias`` (:497-503),
``normalizer`` / ``unnormalizer``.  All arithmetic runs in libdiffco_b200: training is one persistent CUDA launch,
scoring one fused launch.

What differs, and why: the reference takes the ground truth from python-fcl / cuRobo / MoveIt on the URDF's meshes — none of
which is part of the hot path or installed here.  The robot is a URDF path / ``collision_interfaces.URDFRobot`` (kinematic
tree compiled into a joint program) or one of this package's ``model`` robots — either way ``fkine`` is fused into the
kernels — and the geometric ground truth is INJECTED through
``gt_check_func(q) -> (B,) {0, 1}`` — an argument the reference's constructor already has (collision_checkers.py:46,87-88).
"""
from __future__ import annotations

import time
from typing import Callable, Optional, Union

import torch

from . import kernel
from .kernel_perceptrons import DiffCo


class _RobotAdapter:
    """The few members of the reference's ``RobotInterfaceBase`` the checkers touch (``rand_configs``, ``joint_limits``,
    ``_n_dofs``), on top of a ``diffco_b200.model`` robot (``dof``, ``limits``, ``fkine``)."""

    def __init__(self, robot, device):
        self.model = robot
        self._n_dofs = int(robot.dof)
        self.device = torch.device(device)
        self.joint_limits = robot.limits.to(self.device)

    def rand_configs(self, num_cfgs: int) -> torch.Tensor:
        lo, hi = self.joint_limits[:, 0], self.joint_limits[:, 1]
        return torch.rand(num_cfgs, self._n_dofs, device=self.device, dtype=lo.dtype) * (hi - lo) + lo

    def __getattr__(self, name):
        return getattr(self.model, name)


class CollisionChecker:
    """collision_checkers.py:28-125.  ``robot``: a ``diffco_b200.model`` robot; ``gt_check_func``: the geometric ground
    truth (the reference derives it from FCL when none is given; here it has to be given)."""

    def __init__(self, robot=None, robot_base_transform=None, environment=None, robot_topic=None, planning_scene_topic=None,
                 gt_check_func: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, device="cuda") -> None:
        if robot_topic is not None or planning_scene_topic is not None:
            raise NotImplementedError("ROS / MoveIt robots are outside this package: pass a URDF path, a URDFRobot or a "
                                      "diffco_b200.model robot")
        if isinstance(robot, str):  # collision_checkers.py:49-59: a path to a URDF file
            import os

            from .collision_interfaces import URDFRobot

            if not os.path.isfile(robot):
                raise ValueError("Invalid robot URDF file path")
            robot = URDFRobot(robot, name=os.path.basename(robot).split(".")[0], base_transform=robot_base_transform, device=device)
        if robot is None or not hasattr(robot, "fkine") or not hasattr(robot, "limits"):
            raise TypeError("robot must be a URDF path, a URDFRobot or a diffco_b200.model robot (dof, limits, fkine)")
        if gt_check_func is None:
            raise NotImplementedError("the geometric ground truth (python-fcl in the reference) is not part of this package: "
                                      "pass gt_check_func(q) -> (B,) labels in {0, 1}")
        if environment is not None and not callable(gt_check_func):
            raise NotImplementedError("ShapeEnv / PCDEnv environments need python-fcl")
        self.device = torch.device(device)
        self.robot = robot if isinstance(robot, _RobotAdapter) else _RobotAdapter(robot, self.device)
        self.environment = environment
        self.gt_check_func = gt_check_func

    def collision(self, q):
        return self.gt_check_func(q)

    def fkine(self, q, return_collision=False, **kwargs):
        return self.robot.model.fkine(q)

    def normalizer(self, unnormalized_q):
        raise NotImplementedError

    def unnormalizer(self, normalized_q):
        raise NotImplementedError

    def _generate_dataset(self, q, labels, dists, num_samples, fix_joints=None, fix_joint_values=None, verbose=False):
        """collision_checkers.py:108-125."""
        if q is None:
            q = self.robot.rand_configs(num_samples)
        if fix_joints is not None:
            q[:, fix_joints] = torch.tensor(fix_joint_values, dtype=q.dtype, device=q.device)
        num_samples = len(q)
        if labels is None:
            if verbose:
                print("Generating labels...")
                start_time = time.time()
            labels = self.gt_check_func(q)
            if verbose:
                print(f"Labels generated in {time.time() - start_time:.2f}s")
        else:
            labels = (labels > 0).type(q.dtype)
        if dists is None:
            dists = torch.zeros(num_samples, dtype=q.dtype, device=q.device)
        return q, labels.to(q.device), dists


class RBFDiffCo(CollisionChecker):
    """collision_checkers.py:127-316 — the perceptron on raw configurations (no feature map).  (In the reference this
    class leaves ``safety_bias`` / ``perceptron_trained`` / ``_calculate_safety_bias`` to its subclass and cannot be used
    on its own; they are provided here.)"""

    def __init__(self, robot=None, robot_base_transform=None, environment=None, robot_topic=None, planning_scene_topic=None,
                 gt_check_func=None, device="cuda", kernel_func=None, perceptron_class=DiffCo, **perceptron_kwargs) -> None:
        CollisionChecker.__init__(self, robot=robot, robot_base_transform=robot_base_transform, environment=environment,
                                  robot_topic=robot_topic, planning_scene_topic=planning_scene_topic,
                                  gt_check_func=gt_check_func, device=device)
        gamma = perceptron_kwargs.pop("gamma", 10)
        self.kernel_func = kernel.RQKernel(gamma) if kernel_func is None else kernel_func
        self.perceptron = perceptron_class(kernel_func=self.kernel_func, **self._perceptron_kwargs(perceptron_kwargs))
        self.q_verify = None
        self.labels_verify = None
        self.safety_bias = 0
        self.perceptron_trained = False

    def _perceptron_kwargs(self, kwargs):
        return kwargs

    # ------------------------------------------------------------------ fit / update / verify
    def fit(self, q=None, labels=None, dists=None, update=False, exist_mask=None, num_samples=5000, verify_ratio=0.1,
            verbose=False, **get_dataset_kwargs):
        """collision_checkers.py:163-218.  Returns (verify_acc, verify_tpr, verify_tnr) — None each without verification."""
        get_dataset_kwargs["verbose"] = not self.perceptron_trained and verbose
        q, labels, dists = self._generate_dataset(q, labels, dists, num_samples, **get_dataset_kwargs)
        num_samples = len(q)
        labels = (2 * labels - 1).type(q.dtype)
        labels_verify = None
        if 0 < verify_ratio < 1:
            num_verify = int(verify_ratio * num_samples)
            verify_indices = torch.randperm(len(q))[:num_verify]
            verify_mask = torch.zeros(len(q), dtype=torch.bool)
            verify_mask[verify_indices] = True
            verify_mask = verify_mask.to(q.device)
            q_train, q_verify = q[~verify_mask], q[verify_mask]
            labels_train, labels_verify = labels[~verify_mask], labels[verify_mask]
            dists_train = dists[~verify_mask]
            if exist_mask is not None:
                exist_mask = exist_mask[~verify_mask]
        elif verify_ratio:
            raise ValueError(f"verify_ratio should be in (0, 1), got {verify_ratio}")
        else:
            q_train, labels_train, dists_train = q, labels, dists
            q_verify = self.robot.rand_configs(100)
        self.perceptron.train(q_train, labels_train, update=update, exist_mask=exist_mask, max_iteration=len(q_train),
                              distance=dists_train, verbose=verbose)
        self.perceptron.fit_poly(kernel_func=kernel.Polyharmonic(k=1, epsilon=1), target="label")
        self.safety_bias = self._calculate_safety_bias(q_verify)
        if verify_ratio:  # verification needs self.safety_bias
            verify_acc, verify_tpr, verify_tnr = self.verify(q_verify, labels_verify, verbose=verbose)
            self.q_verify = q_verify
        else:
            verify_acc, verify_tpr, verify_tnr = None, None, None
        self.perceptron_trained = True
        return verify_acc, verify_tpr, verify_tnr

    def update(self, q=None, labels=None, dists=None, exploit_std=0.3, num_samples=100, num_exploit_samples=None,
               num_explore_samples=None, verify=False, verbose=False):
        """collision_checkers.py:220-252: new samples around the support points (exploit) and uniform ones (explore), the old
        supports appended and marked in ``exist_mask``; training restarts from the current model (jump start)."""
        num_exploit_samples = num_samples if num_exploit_samples is None else num_exploit_samples
        num_explore_samples = num_samples if num_explore_samples is None else num_explore_samples
        exist_mask = None
        if q is None:
            sp = self.perceptron.support_points
            if num_exploit_samples > len(sp):
                mul = (num_exploit_samples // len(sp)) + (num_exploit_samples % len(sp) > 0)
                selected = torch.arange(len(sp))
            else:
                mul = 1
                selected = torch.randperm(len(sp))[:num_exploit_samples]
            chosen = sp[selected.to(sp.device)]
            device, dtype = chosen.device, chosen.dtype
            lim = self.robot.joint_limits.to(device=device, dtype=dtype)
            exploit = torch.randn(mul, len(chosen), self.robot._n_dofs, dtype=dtype, device=device) * exploit_std + chosen[None]
            exploit = torch.clamp(exploit, min=lim[:, 0], max=lim[:, 1]).reshape(-1, self.robot._n_dofs)
            explore = self.robot.rand_configs(num_explore_samples).to(device=device, dtype=dtype)
            q = torch.cat([exploit, explore, sp], dim=0)
            exist_mask = torch.zeros(len(q), dtype=torch.bool, device=device)
            exist_mask[-len(sp):] = True
        return self.fit(q, labels, dists, update=True, exist_mask=exist_mask, verify_ratio=verify, verbose=verbose)

    def verify(self, q_verify=None, labels_verify=None, num_samples=None, verbose=False):
        """collision_checkers.py:254-290.  Returns the BIASED (acc, tpr, tnr) like the reference; prints both."""
        if q_verify is None:
            if num_samples is not None:
                q_verify = self.robot.rand_configs(num_samples)
                self.q_verify = q_verify
            elif self.q_verify is not None:
                q_verify = self.q_verify
            else:
                raise ValueError("self.q_verify or num_samples should be provided")
        scores_verify = self.perceptron.poly_score(q_verify)
        preds_verify = 2 * (scores_verify > 0) - 1
        biased_preds_verify = 2 * (scores_verify + self.safety_bias > 0) - 1
        if labels_verify is None:
            labels_verify = self.gt_check_func(q_verify)
            labels_verify = (2 * labels_verify - 1).type(q_verify.dtype)
        labels_verify = labels_verify.to(preds_verify.device)
        preds_verify = preds_verify.reshape_as(labels_verify)
        biased_preds_verify = biased_preds_verify.reshape_as(labels_verify)
        n_total, n_pos, n_neg = len(preds_verify), (labels_verify == 1).sum(), (labels_verify == -1).sum()

        def rates(pred):
            acc = torch.sum(pred == labels_verify, dtype=torch.float32) / n_total
            tpr = torch.sum(pred[labels_verify == 1] == 1, dtype=torch.float32) / n_pos
            tnr = torch.sum(pred[labels_verify == -1] == -1, dtype=torch.float32) / n_neg
            return acc, tpr, tnr

        test_acc, test_tpr, test_tnr = rates(preds_verify)
        if verbose:
            print(f"Positive labels: {n_pos.item()}, Negative labels: {n_neg.item()}")
            print(f"Test acc: {test_acc:.4f}, TPR {test_tpr:.4f}, TNR {test_tnr:.4f}")
        self.last_unbiased_rates = (test_acc, test_tpr, test_tnr)
        test_acc, test_tpr, test_tnr = rates(biased_preds_verify)
        if verbose:
            print(f"Biased Test acc: {test_acc:.4f}, TPR {test_tpr:.4f}, TNR {test_tnr:.4f}")
        return test_acc, test_tpr, test_tnr

    # ------------------------------------------------------------------ scoring
    def collision(self, q):
        return self.collision_score(q) > 0

    def collision_score(self, q, bias: Union[float, torch.Tensor, None] = None):
        """collision_checkers.py:295-303 — q of shape (..., num_dof) -> (..., 1) scores + bias."""
        bias = self.safety_bias if bias is None else bias
        shape_q = q.shape
        raw = self.perceptron.poly_score(q.reshape(-1, shape_q[-1]))
        raw = raw.reshape(shape_q[:-1] + raw.shape[1:])
        return raw + bias

    def _calculate_safety_bias(self, q_verify):
        """collision_checkers.py:497-503: a third of the smaller of |min score| and |max score| over the verification set."""
        scores = self.perceptron.poly_score(q_verify)[:, 0]
        min_score, max_score = scores.min(), scores.max()
        return min(min_score.abs(), max_score.abs()) / 3

    def normalizer(self, unnormalized_q):
        lim = self.robot.joint_limits.to(unnormalized_q.device)
        return (unnormalized_q - lim[:, 0]) / (lim[:, 1] - lim[:, 0])

    def unnormalizer(self, normalized_q):
        lim = self.robot.joint_limits.to(normalized_q.device)
        return normalized_q * (lim[:, 1] - lim[:, 0]) + lim[:, 0]


class ForwardKinematicsDiffCo(RBFDiffCo):
    """collision_checkers.py:318-509 — the perceptron on forward-kinematics control points (``transform = robot.fkine``, fused
    into the CUDA kernels).  ``collision_score`` also takes pre-computed link positions (``q_link_pos``)."""

    def _perceptron_kwargs(self, kwargs):
        return dict(kwargs, transform=self.robot.model.fkine)

    def tensorized_fkine(self, q, return_collision=False):
        return self.robot.model.fkine(q)

    def _uniform_sample_on_transformed_manifold(self, transform, num_samples):
        """collision_checkers.py:396-453: rejection sampling with acceptance ~ sqrt(det(J J^T + 1e-4 I)), J the Jacobian of the
        feature map — uniform on the image manifold instead of uniform in configuration space."""

        def jac_det(q):
            q = q.clone().detach().requires_grad_(True)
            pos = transform(q).reshape(q.shape[0], -1)
            rows = []
            for i in range(pos.shape[1]):  # the reference's row-by-row backward (its comment: functorch does not apply here)
                (g,) = torch.autograd.grad(pos[:, i].sum(), q, retain_graph=True)
                rows.append(g)
            jac = torch.stack(rows, dim=1)  # (B, F, D)
            if jac.shape[-2] > jac.shape[-1]:
                jac = jac.transpose(-2, -1)
            eye = torch.eye(jac.shape[-2], device=jac.device, dtype=jac.dtype)
            return torch.linalg.det(jac @ jac.transpose(-2, -1) + 1e-4 * eye).sqrt()

        rand_q = self.robot.rand_configs(num_samples)
        det = jac_det(rand_q)
        max_det = 1.1 * det.max()
        valid, cnt = [], 0
        while True:
            accept = det > torch.rand(len(rand_q), device=rand_q.device, dtype=rand_q.dtype) * max_det
            valid.append(rand_q[accept])
            cnt += int(accept.sum())
            if cnt >= num_samples:
                break
            rand_q = self.robot.rand_configs(num_samples)
            det = jac_det(rand_q)
        return torch.cat(valid, dim=0)[:num_samples]

    def _generate_dataset(self, q, labels, dists, num_samples, verbose=False, sample_transform=None, **kwargs):
        """collision_checkers.py:455-473."""
        transform = None
        if sample_transform == "fkine":
            transform = self.tensorized_fkine
        elif callable(sample_transform):
            transform = sample_transform
        elif sample_transform is not None:
            raise ValueError(f"Invalid sample_transform: {sample_transform}")
        if transform is not None:
            q = self._uniform_sample_on_transformed_manifold(transform, num_samples)
            num_samples = len(q)
        return super()._generate_dataset(q, labels, dists, num_samples, verbose=verbose, **kwargs)

    def collision_score(self, q: Optional[torch.Tensor] = None, bias: Union[float, torch.Tensor, None] = None,
                        q_link_pos: Optional[torch.Tensor] = None):
        """collision_checkers.py:475-495 — q (..., num_dof) or q_link_pos (..., num_links, dim)."""
        bias = self.safety_bias if bias is None else bias
        if q is not None:
            shape_q = q.shape
            raw = self.perceptron.poly_score(point=q.reshape(-1, shape_q[-1]))
            raw = raw.reshape(shape_q[:-1] + raw.shape[1:])
        elif q_link_pos is not None:
            shape = q_link_pos.shape
            raw = self.perceptron.poly_score(transformed_point=q_link_pos.reshape(-1, *shape[-2:]))
            raw = raw.reshape(shape[:-2] + raw.shape[1:])
        else:
            raise ValueError("q or q_link_pos must be given")
        return raw + bias
