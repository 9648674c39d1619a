"""URDF robots — host-side mirror of ``URDFRobot`` (collision_interfaces/urdf_interface.py:330-553 of the reference) for
the kinematics only.

The reference loads the file with yourdfpy, wraps every link in a ``RigidBody`` node (rigid_body.py:19-81) and evaluates
forward kinematics by recursing over the tree in Python, one small torch op per joint (rigid_body.py:86-141).  Here the
file is parsed with ``xml.etree`` (no yourdfpy / trimesh needed), the tree is flattened parents-first into
``dc_fk_desc.tree`` (a *joint program*: fixed transform, joint kind, driving column of q, mimic law, output slot) and the
CUDA side runs it per query inside the fused score kernel (``DC_FK_JOINT_TREE``, csrc/dc_fk.cuh) — forward and the
analytic J^T product.  ``compute_forward_kinematics_all_links`` returns the reference's dictionary of per-link
(translation, rotation) from ``dc_fk_tree_frames``.

Constants follow the reference: joint origins, axes and limits are rounded to float32 (urdf_interface.py:583-600), the
fixed rotation is Rz(yaw) Ry(pitch) Rx(roll) (rigid_body.py:96-99), a revolute axis is matched to +-x / +-y, otherwise
+-z (rigid_body.py:103-108), joints are ordered by the order of the links in the file and mimic joints share their
master's column (urdf_interface.py:377-388,536-540).

Collision checking against meshes (``collision``; python-fcl in the reference) is outside this package: pass the ground
truth to the checkers through ``gt_check_func``.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib, functional
from ..model import Model


def _floats(text, n, default):
    if text is None:
        return list(default)
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return vals


def _origin(el):
    o = el.find("origin") if el is not None else None
    if o is None:
        return [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    return _floats(o.get("xyz"), 3, (0, 0, 0)), _floats(o.get("rpy"), 3, (0, 0, 0))


def _rpy_matrix(rpy) -> np.ndarray:
    """Rz(yaw) Ry(pitch) Rx(roll) in float64 (rigid_body.py:96-99)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return (rz @ ry) @ rx


def parse_urdf(source: str) -> Tuple[List[dict], Dict[str, dict]]:
    """Links (file order) and joints of a URDF file (``source``: a path, or the XML text itself).  Each link:
    {name, collision_origins: [(xyz, rpy)]}; each joint: {name, type, parent, child, xyz, rpy, axis, lower, upper, mimic}."""
    root = ET.fromstring(source) if source.lstrip().startswith("<") else ET.parse(source).getroot()
    if root.tag != "robot":
        raise ValueError("not a URDF document (root element must be <robot>)")
    links = []
    for el in root.findall("link"):
        links.append({"name": el.get("name"), "collision_origins": [_origin(c) for c in el.findall("collision")]})
    joints = {}
    for el in root.findall("joint"):
        xyz, rpy = _origin(el)
        ax = el.find("axis")
        lim = el.find("limit")
        mim = el.find("mimic")
        joints[el.get("name")] = {
            "name": el.get("name"), "type": el.get("type"),
            "parent": el.find("parent").get("link"), "child": el.find("child").get("link"),
            "xyz": xyz, "rpy": rpy,
            "axis": _floats(ax.get("xyz") if ax is not None else None, 3, (1, 0, 0)),  # URDF default axis
            "lower": float(lim.get("lower", 0.0)) if lim is not None else None,
            "upper": float(lim.get("upper", 0.0)) if lim is not None else None,
            "mimic": None if mim is None else {"joint": mim.get("joint"), "multiplier": float(mim.get("multiplier", 1.0)),
                                               "offset": float(mim.get("offset", 0.0))},
        }
    return links, joints


class _Body:
    """What the checkers read from the reference's ``RigidBody`` nodes: name, joint_type, joint_trans(), dof_idx."""

    def __init__(self, name, joint_name, joint_type, trans, rpy, axis, limits, mimic):
        self.name, self.joint_name, self.joint_type = name, joint_name, joint_type
        self._trans = torch.tensor(trans, dtype=torch.float32).reshape(1, 3)
        self._rpy = torch.tensor(rpy, dtype=torch.float32).reshape(1, 3)
        self.joint_axis = torch.tensor(axis, dtype=torch.float32).reshape(1, 3)
        self.joint_limits = limits
        self.joint_mimic = mimic
        self.dof_idx = None
        self.parent_idx = -1
        self.collision_origins = []

    def joint_trans(self):
        return self._trans

    def joint_rot_angles(self):
        return self._rpy


class URDFRobot(Model):
    """urdf_interface.py:330-553.  ``fkine(q)`` — the feature map of ``ForwardKinematicsDiffCo`` (collision_checkers.py:345-391):
    origins of the link frames whose joint has a non-zero translation, (B, L, 3) — is fused into the score kernels through
    ``fk_desc``; the reference stacks the same points as (B, 3, L) (``tensorized_fkine``), which is the same feature set."""

    def __init__(self, urdf_path, name="", base_transform: Optional[torch.Tensor] = None, device="cpu", setup_acm=False,
                 load_visual_meshes=False):
        if load_visual_meshes:
            raise NotImplementedError("meshes (trimesh) are outside this package")
        self.name = name
        self._device = torch.device(device)
        links, joints = parse_urdf(urdf_path)
        child_joint = {j["child"]: j for j in joints.values()}
        self._bodies: List[_Body] = []
        self._body_name_to_idx_map: Dict[str, int] = {}
        self._n_dofs = 0
        self._controlled_joints: List[int] = []
        self._mimic_joints = defaultdict(list)
        for idx, link in enumerate(links):
            j = child_joint.get(link["name"])
            if j is None:
                body = _Body(link["name"], "base_joint", "fixed", (0, 0, 0), (0, 0, 0), (0, 0, 0), None, None)
            else:
                if j["type"] not in ("fixed", "revolute", "continuous", "prismatic"):
                    raise NotImplementedError(f"joint {j['name']}: type {j['type']!r} (the reference evaluates fixed, revolute, "
                                              "continuous and prismatic joints, rigid_body.py:101-124)")
                limits = None
                if j["type"] != "fixed" and j["lower"] is not None:
                    limits = {"lower": j["lower"], "upper": j["upper"]}
                body = _Body(link["name"], j["name"], j["type"], j["xyz"], j["rpy"],
                             j["axis"] if j["type"] != "fixed" else (0, 0, 0), limits, j["mimic"])
            body.collision_origins = link["collision_origins"]
            if body.joint_type != "fixed":
                if body.joint_mimic is None:
                    body.dof_idx = self._n_dofs
                    self._n_dofs += 1
                    self._controlled_joints.append(idx)
                else:
                    self._mimic_joints[joints[body.joint_mimic["joint"]]["child"]].append(body.name)
            self._bodies.append(body)
            self._body_name_to_idx_map[body.name] = idx
        for body in self._bodies:
            if body.joint_name != "base_joint":
                body.parent_idx = self._body_name_to_idx_map[joints[body.joint_name]["parent"]]

        # joint limits (urdf_interface.py:405-419)
        self.joint_limits = torch.zeros((self._n_dofs, 2))
        for i, bi in enumerate(self._controlled_joints):
            b = self._bodies[bi]
            if b.joint_type in ("revolute", "prismatic"):
                lo, hi = (b.joint_limits["lower"], b.joint_limits["upper"]) if b.joint_limits is not None else (-np.pi, np.pi)
            else:
                lo, hi = -2 * np.pi, 2 * np.pi
            self.joint_limits[i, 0], self.joint_limits[i, 1] = lo, hi
        self.dof = self._n_dofs
        self.limits = self.joint_limits

        base = torch.eye(4, dtype=torch.float64) if base_transform is None else torch.as_tensor(base_transform).double().cpu()
        self.base_transform = base
        # features of ForwardKinematicsDiffCo: links whose joint translation is non-zero (collision_checkers.py:356-358)
        self.unique_position_link_names = [b.name for b in self._bodies if bool(torch.any(b.joint_trans() != 0))]
        self._compile(base)
        self._finalize()

    # ------------------------------------------------------------------ joint program
    def _compile(self, base):
        master_col = {self._bodies[bi].name: i for i, bi in enumerate(self._controlled_joints)}
        for master, followers in self._mimic_joints.items():
            for f in followers:
                if master in master_col:
                    master_col[f] = master_col[master]
        order, stack = [], [0]  # the reference recurses from the first link of the file (urdf_interface.py:543)
        children = defaultdict(list)
        for i, b in enumerate(self._bodies):
            if b.parent_idx >= 0:
                children[b.parent_idx].append(i)
        while stack:
            i = stack.pop()
            order.append(i)
            stack.extend(reversed(children[i]))
        if len(order) > _lib.DC_MAX_TREE_NODES:
            raise ValueError(f"{len(order)} bodies; the joint program holds {_lib.DC_MAX_TREE_NODES}")
        if self._n_dofs > _lib.DC_MAX_DOF:
            raise ValueError(f"{self._n_dofs} joints; the kernels hold {_lib.DC_MAX_DOF}")
        node_of = {bi: n for n, bi in enumerate(order)}
        slots = {name: s for s, name in enumerate(self.unique_position_link_names)}
        missing = [n for n in slots if self._body_name_to_idx_map[n] not in node_of]
        if missing:
            raise ValueError(f"links not connected to the first link of the file: {missing}")
        if 3 * len(slots) > _lib.DC_MAX_FEATURES or not slots:
            raise ValueError(f"{len(slots)} feature links; the kernels hold 1..{_lib.DC_MAX_FEATURES // 3}")
        d = _lib.FkDesc()
        d.type = _lib.DC_FK_JOINT_TREE
        d.dof, d.n_points, d.point_dim, d.n_nodes = self._n_dofs, len(slots), 3, len(order)
        self.node_names = []
        for n, bi in enumerate(order):
            b = self._bodies[bi]
            nd = d.tree[n]
            nd.parent = node_of[b.parent_idx] if b.parent_idx >= 0 else -1
            nd.out_slot = slots.get(b.name, -1)
            rot = _rpy_matrix(b.joint_rot_angles()[0].double().tolist())
            trans = b.joint_trans()[0].double().numpy()
            if nd.parent < 0:  # fold the robot's base transform into the root
                rot, trans = base[:3, :3].numpy() @ rot, base[:3, :3].numpy() @ trans + base[:3, 3].numpy()
            nd.rot[:] = rot.reshape(-1).tolist()
            nd.trans[:] = trans.tolist()
            nd.mimic_mul, nd.mimic_off = 1.0, 0.0
            nd.q_index = master_col.get(b.name, -1) if b.joint_type != "fixed" else -1
            ax = b.joint_axis[0].double().tolist()
            if b.joint_type == "fixed" or nd.q_index < 0:
                nd.joint, nd.q_index = _lib.DC_JOINT_FIXED, -1
            elif b.joint_type == "prismatic":
                nd.joint = _lib.DC_JOINT_PRISMATIC
                nd.axis[:] = ax
            else:  # rigid_body.py:103-108
                if abs(ax[0]) == 1:
                    nd.joint, sgn = _lib.DC_JOINT_REV_X, np.sign(ax[0])
                elif abs(ax[1]) == 1:
                    nd.joint, sgn = _lib.DC_JOINT_REV_Y, np.sign(ax[1])
                else:
                    nd.joint, sgn = _lib.DC_JOINT_REV_Z, np.sign(ax[2])
                nd.axis[:] = [float(sgn), 0.0, 0.0]
            if b.joint_mimic is not None and nd.q_index >= 0:
                nd.mimic_mul, nd.mimic_off = b.joint_mimic["multiplier"], b.joint_mimic["offset"]
            self.node_names.append(b.name)
        self.fk_desc = d

    # ------------------------------------------------------------------ reference protocol
    def rand_configs(self, num_cfgs):
        lim = self.joint_limits.to(self._device)
        return torch.rand(num_cfgs, self._n_dofs, device=self._device) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]

    def collision(self, q, other=None, show=False):
        raise NotImplementedError("mesh collision checking (python-fcl in the reference) is outside this package: give the "
                                  "checkers the ground truth through gt_check_func")

    def compute_forward_kinematics_all_links(self, q, return_collision=False):
        """urdf_interface.py:517-553: {link: [(translation (B, 3), rotation (B, 3, 3)), ...]} — one pose per link, or one per
        collision geometry of the link with ``return_collision`` (links without collision geometry then have none)."""
        q = torch.as_tensor(q)
        if q.ndim == 1:
            q = q[None]
        fr = functional.fk_tree_frames(self.fk_desc, q)  # (B, n_nodes, 12)
        rot, trans = fr[..., :9].reshape(len(q), -1, 3, 3), fr[..., 9:]
        out = {}
        for n, name in enumerate(self.node_names):
            if not return_collision:
                out[name] = [(trans[:, n], rot[:, n])]
                continue
            poses = []
            for xyz, rpy in self._bodies[self._body_name_to_idx_map[name]].collision_origins:
                ro = torch.as_tensor(_rpy_matrix(np.float32(rpy)), dtype=rot.dtype, device=rot.device)
                to = torch.as_tensor(np.float32(xyz), dtype=rot.dtype, device=rot.device)
                poses.append((trans[:, n] + rot[:, n] @ to, rot[:, n] @ ro))
            out[name] = poses
        return out


class MultiURDFRobot(Model):
    """urdf_interface.py:700-870: several URDF robots side by side.  Configurations are the robots' configurations
    concatenated (``split_configs``); the feature map of ``ForwardKinematicsDiffCo`` is the concatenation of the robots'
    feature links in robot order (collision_checkers.py:347-351,374-384).  The joint programs are merged into ONE
    ``dc_fk_desc`` (node, column and output-slot indices shifted), so the combined map is fused into the score kernels like
    a single robot — within the limits of the descriptor (24 bodies, 16 joints, 21 feature links in total)."""

    def __init__(self, urdf_robots: Optional[List[URDFRobot]] = None, urdf_paths: Optional[List[str]] = None,
                 names: Optional[List[str]] = None, base_transforms: Optional[List[torch.Tensor]] = None, name: str = None,
                 device="cpu", setup_acm=False, load_visual_meshes=False):
        if urdf_robots is not None:
            self.urdf_robots = list(urdf_robots)
            if len({r.name for r in self.urdf_robots}) != len(self.urdf_robots):
                raise AssertionError("Robot names must be unique")
            self.name = "_".join(r.name for r in self.urdf_robots) if name is None else name
            self._device = self.urdf_robots[0]._device
        else:
            base_transforms = base_transforms if base_transforms is not None else [None] * len(urdf_paths)
            self.urdf_robots = [URDFRobot(p, name=n, base_transform=b, device=device, setup_acm=setup_acm,
                                          load_visual_meshes=load_visual_meshes)
                                for p, n, b in zip(urdf_paths, names, base_transforms)]
            self.name = "_".join(names) if name is None else name
            self._device = torch.device(device)
        self.inter_robot_acm = None
        self._bodies = [list(r._bodies) for r in self.urdf_robots]
        self._n_dofs = sum(r._n_dofs for r in self.urdf_robots)
        self.dof = self._n_dofs
        self.joint_limits = torch.cat([r.joint_limits for r in self.urdf_robots], dim=0)
        self.limits = self.joint_limits
        self.unique_position_link_names = [(i, n) for i, r in enumerate(self.urdf_robots) for n in r.unique_position_link_names]
        self._merge()
        self._finalize()

    def _merge(self):
        n_nodes = sum(r.fk_desc.n_nodes for r in self.urdf_robots)
        n_slots = sum(r.fk_desc.n_points for r in self.urdf_robots)
        if n_nodes > _lib.DC_MAX_TREE_NODES or self._n_dofs > _lib.DC_MAX_DOF or 3 * n_slots > _lib.DC_MAX_FEATURES:
            raise ValueError(f"{n_nodes} bodies / {self._n_dofs} joints / {n_slots} feature links in total; the joint program holds "
                             f"{_lib.DC_MAX_TREE_NODES} / {_lib.DC_MAX_DOF} / {_lib.DC_MAX_FEATURES // 3}")
        d = _lib.FkDesc()
        d.type = _lib.DC_FK_JOINT_TREE
        d.dof, d.n_points, d.point_dim, d.n_nodes = self._n_dofs, n_slots, 3, n_nodes
        node0 = col0 = slot0 = 0
        self.node_names = []
        for i, r in enumerate(self.urdf_robots):
            src = r.fk_desc
            for k in range(src.n_nodes):
                nd = _lib.TreeNode.from_buffer_copy(src.tree[k])
                if nd.parent >= 0:
                    nd.parent += node0
                if nd.q_index >= 0:
                    nd.q_index += col0
                if nd.out_slot >= 0:
                    nd.out_slot += slot0
                d.tree[node0 + k] = nd
                self.node_names.append((i, r.node_names[k]))
            node0, col0, slot0 = node0 + src.n_nodes, col0 + src.dof, slot0 + src.n_points
        self.fk_desc = d

    def rand_configs(self, num_cfgs):
        return torch.cat([r.rand_configs(num_cfgs) for r in self.urdf_robots], dim=1)

    def split_configs(self, q):
        return torch.split(q, [r._n_dofs for r in self.urdf_robots], dim=1)

    def collision(self, q, other=None, show=False):
        raise NotImplementedError("mesh collision checking (python-fcl in the reference) is outside this package: give the "
                                  "checkers the ground truth through gt_check_func")

    def compute_forward_kinematics_all_links(self, q, return_collision=False):
        """One dictionary per robot (urdf_interface.py:857-862)."""
        q = torch.as_tensor(q)
        if q.ndim == 1:
            q = q[None]
        return [r.compute_forward_kinematics_all_links(qi, return_collision=return_collision)
                for r, qi in zip(self.urdf_robots, self.split_configs(q))]

    def update_acm(self, link_pairs):
        self.inter_robot_acm = set(link_pairs) if self.inter_robot_acm is None else self.inter_robot_acm.union(link_pairs)
