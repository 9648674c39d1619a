"""Path / angle helpers — host-side mirror of the pieces of the reference's ``diffco/utils.py`` that the trajectory
optimisers use (the rotation / DH builders of that file live inside the CUDA FK, csrc/dc_fk.cuh).

Everything here is ordinary differentiable torch code on whatever device the waypoints live on: these functions see
W ~ 20 waypoints, not batches.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def wrap2pi(theta):
    """utils.py:51-52 — map angles to [-pi, pi)."""
    return (math.pi + theta) % (2 * math.pi) - math.pi


def se2_wrap2pi(x):
    """utils.py:54-55."""
    return torch.cat([x[..., :2], wrap2pi(x[..., 2:3])], dim=-1)


def anglin(q1, q2, num=50, endpoint=True):
    """utils.py:60-64 — linspace between two angle vectors along the short way round."""
    q1 = torch.as_tensor(q1, dtype=torch.float32)
    q2 = torch.as_tensor(q2, dtype=torch.float32)
    span = wrap2pi(q2 - q1).numpy()
    dq = torch.from_numpy(np.linspace(np.zeros_like(span), span, num, endpoint))
    return wrap2pi(q1 + dq)


def make_continue(q, max_gap=math.pi):
    """utils.py:79-85 — undo 2*pi jumps along a sequence of angles."""
    q = torch.as_tensor(q, dtype=torch.float32)
    jump = torch.zeros_like(q)
    step = q[1:] - q[:-1]
    jump[1:] = (step.abs() > max_gap) * torch.sign(step)
    return q - torch.cumsum(jump, dim=0) * 2 * math.pi


def segment_steps(q, max_step=2.0, max_step_num=None):
    """Number of interpolation points per segment used by dense_path: ceil(|q[i+1]-q[i]| / max_step)."""
    with torch.no_grad():
        seg = torch.norm(q[1:] - q[:-1], dim=-1)
        if max_step_num is not None:
            max_step = max(max_step, seg.sum().item() / max_step_num)
        return torch.ceil(seg / max_step).to(torch.long), max_step


def dense_path(q, max_step=2.0, max_step_num=None):
    """utils.py:87-102 — per segment ceil(|dq|/max_step) points q[i] + k*max_step*dq/|dq| (k = 0, 1, ...), then the last
    waypoint; differentiable w.r.t. q (the optimisers backpropagate through the normalised direction)."""
    steps, max_step = segment_steps(q, max_step, max_step_num)
    seg_index = torch.repeat_interleave(torch.arange(len(q) - 1, device=q.device), steps.to(q.device))
    first = torch.cumsum(steps, 0) - steps
    k = (torch.arange(len(seg_index), device=q.device) - first.to(q.device)[seg_index]).to(q.dtype).reshape(-1, 1)
    delta = q[1:] - q[:-1]
    dist = delta.norm(dim=-1, keepdim=True)
    # a zero-length segment contributes no points (steps == 0): keep its 0/0 direction out of the autograd graph
    # (the reference's per-segment loop never evaluates it, utils.py:93-99)
    unit = delta * max_step / torch.where(dist > 0, dist, torch.ones_like(dist))
    pts = q[:-1][seg_index] + k * unit[seg_index]
    dense = torch.cat([pts, q[-1:]])
    assert torch.all(dense[0] == q[0]) and torch.all(dense[-1] == q[-1])
    return dense
