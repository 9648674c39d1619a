"""CPU tests of the URDF front end (diffco_b200/collision_interfaces/urdf_interface.py): parsing, the reference's
conventions for joint order / mimic joints / limits / feature links, and the compiled joint program — no GPU needed (the
descriptor is plain host data; the arithmetic on it is covered by tests/test_oracle_vs_golden.py and tests/test_gpu_urdf.py)."""
import os

import numpy as np
import pytest
import torch

from diffco_b200 import _lib
from diffco_b200.collision_interfaces import URDFRobot, parse_urdf

DATA = os.path.join(os.path.dirname(__file__), "data")


def test_parse_and_compile_arm_with_gripper():
    robot = URDFRobot(os.path.join(DATA, "arm7_gripper.urdf"))
    # controlled joints follow the order of the LINKS in the file (urdf_interface.py:377-388): `tool` is listed second but
    # hangs off a fixed joint; the mimic finger shares its master's column
    assert robot._n_dofs == 8 and robot.dof == 8
    assert [robot._bodies[i].name for i in robot._controlled_joints] == ["l1", "l2", "l3", "l4", "l5", "l6", "l7", "finger_a"]
    assert dict(robot._mimic_joints) == {"finger_a": ["finger_b"]}
    lim = robot.joint_limits
    assert torch.allclose(lim[0], torch.tensor([-2.9, 2.9])) and torch.allclose(lim[4], torch.tensor([-2 * np.pi, 2 * np.pi]))
    assert torch.allclose(lim[7], torch.tensor([0.0, 0.04]))
    # feature links: those whose joint origin has a translation (collision_checkers.py:356-358) — l2, l6 and tool do not
    assert robot.unique_position_link_names == ["l1", "l3", "l4", "l5", "l7", "flange", "finger_a", "finger_b"]
    d = robot.fk_desc
    assert (d.type, d.dof, d.n_points, d.point_dim, d.n_nodes) == (_lib.DC_FK_JOINT_TREE, 8, 8, 3, 12)
    nodes = {name: d.tree[i] for i, name in enumerate(robot.node_names)}
    assert robot.node_names[0] == "base" and nodes["base"].parent == -1 and nodes["base"].joint == _lib.DC_JOINT_FIXED
    for i, name in enumerate(robot.node_names[1:], 1):
        assert 0 <= d.tree[i].parent < i  # parents first
    assert nodes["l2"].joint == _lib.DC_JOINT_REV_Z and nodes["l2"].axis[0] == -1.0      # axis 0 0 -1
    assert nodes["l3"].joint == _lib.DC_JOINT_REV_Y and nodes["l4"].joint == _lib.DC_JOINT_REV_X and nodes["l4"].axis[0] == -1.0
    assert nodes["finger_b"].joint == _lib.DC_JOINT_PRISMATIC and nodes["finger_b"].q_index == nodes["finger_a"].q_index == 7
    assert (nodes["finger_b"].mimic_mul, nodes["finger_b"].mimic_off) == (0.5, 0.01)
    assert nodes["flange"].q_index == -1 and nodes["flange"].out_slot == 5 and nodes["l2"].out_slot == -1
    rot = np.array(list(nodes["l2"].rot)).reshape(3, 3)  # rpy = (-pi/2, 0, 0) rounded to float32
    assert np.allclose(rot, [[1, 0, 0], [0, 0, 1], [0, -1, 0]], atol=1e-6)
    cfg = robot.rand_configs(5)
    assert cfg.shape == (5, 8)


def test_base_transform_defaults_and_odd_axes():
    base = torch.tensor([[0.0, -1.0, 0.0, 0.3], [1.0, 0.0, 0.0, -0.2], [0.0, 0.0, 1.0, 0.1], [0, 0, 0, 1.0]])
    robot = URDFRobot(os.path.join(DATA, "torso_two_arms.urdf"), base_transform=base)
    d = robot.fk_desc
    root = d.tree[0]
    assert np.allclose(np.array(list(root.rot)).reshape(3, 3), base[:3, :3].numpy()) and np.allclose(list(root.trans), base[:3, 3].numpy())
    nodes = {name: d.tree[i] for i, name in enumerate(robot.node_names)}
    # an axis that is no coordinate axis: the reference rotates about sign(z) z (rigid_body.py:103-108)
    assert nodes["left_3"].joint == _lib.DC_JOINT_REV_Z and nodes["left_3"].axis[0] == -1.0
    # a revolute joint without <limit>: +-pi (urdf_interface.py:411-416)
    i = [robot._bodies[b].name for b in robot._controlled_joints].index("right_2")
    assert torch.allclose(robot.joint_limits[i], torch.tensor([-np.pi, np.pi]))
    assert nodes["torso"].joint == _lib.DC_JOINT_PRISMATIC and np.allclose(list(nodes["torso"].axis), np.float32([0.1, 0, 0.9]))


def test_rejects_what_the_joint_program_cannot_hold():
    xml = "<robot name='r'><link name='a'/><link name='b'/><joint name='j' type='floating'><parent link='a'/><child link='b'/></joint></robot>"
    with pytest.raises(NotImplementedError):
        URDFRobot(xml)
    with pytest.raises(ValueError):
        parse_urdf("<notarobot/>")
    links = "".join(f"<link name='l{i}'/>" for i in range(30))
    joints = "".join(f"<joint name='j{i}' type='fixed'><origin xyz='0 0 0.1'/><parent link='l{i}'/><child link='l{i + 1}'/></joint>"
                     for i in range(29))
    with pytest.raises(ValueError):
        URDFRobot(f"<robot name='long'>{links}{joints}</robot>")  # 30 bodies > DC_MAX_TREE_NODES


def test_multi_robot_merges_the_joint_programs():
    """MultiURDFRobot (urdf_interface.py:700-870): configurations and feature links concatenated in robot order; the merged
    descriptor is the robots' programs with node / column / slot indices shifted."""
    from diffco_b200.collision_interfaces import MultiURDFRobot

    a = URDFRobot(os.path.join(DATA, "torso_two_arms.urdf"), name="t")
    shift = torch.eye(4)
    shift[0, 3] = 1.0
    b = URDFRobot(os.path.join(DATA, "torso_two_arms.urdf"), name="u", base_transform=shift)
    multi = MultiURDFRobot(urdf_robots=[a, b])
    assert multi.name == "t_u" and multi.dof == a.dof + b.dof == 14
    assert multi.unique_position_link_names == [(0, n) for n in a.unique_position_link_names] + [(1, n) for n in b.unique_position_link_names]
    d = multi.fk_desc
    assert (d.n_nodes, d.n_points, d.dof) == (18, 16, 14)
    for k in range(b.fk_desc.n_nodes):
        m, s = d.tree[a.fk_desc.n_nodes + k], b.fk_desc.tree[k]
        assert m.parent == (s.parent + 9 if s.parent >= 0 else -1) and m.q_index == (s.q_index + 7 if s.q_index >= 0 else -1)
        assert m.out_slot == (s.out_slot + 8 if s.out_slot >= 0 else -1) and list(m.rot) == list(s.rot) and list(m.trans) == list(s.trans)
    q = multi.rand_configs(4)
    assert q.shape == (4, 14) and [t.shape[1] for t in multi.split_configs(q)] == [7, 7]
    assert torch.equal(multi.joint_limits, torch.cat([a.joint_limits, b.joint_limits]))
    with pytest.raises(AssertionError):
        MultiURDFRobot(urdf_robots=[a, a])
    with pytest.raises(ValueError):  # 3 x 9 bodies > 24
        MultiURDFRobot(urdf_paths=[os.path.join(DATA, "torso_two_arms.urdf")] * 3, names=["a", "b", "c"])
