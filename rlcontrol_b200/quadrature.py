"""Integration grids of the ForwardKL / ReverseKL agents (host side, built once per network).

The reference builds them with the third-party ``quadpy`` (unpinned, requirements.txt:4;
``quadpy.c1.clenshaw_curtis(N_param)`` forwardkl_network.py:60-71, reversekl_network.py:65-76, and a
Smolyak combination of 1-D Clenshaw-Curtis rules when action_dim > 1, forwardkl_network.py:73-102).
quadpy is not needed here: the rule is the classical closed form

    x_j = -cos(j pi / (n-1)),   w_j = c_j/(n-1) * (1 - sum_{k=1}^{floor((n-1)/2)} b_k/(4k^2-1) cos(2 k j pi/(n-1)))

with c_j = 1 at the two end points and 2 inside, b_k = 1 when 2k = n-1 and 2 otherwise.
"""
from __future__ import annotations

import itertools
from math import comb
from typing import Tuple

import numpy as np


def clenshaw_curtis(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """n-point Clenshaw-Curtis nodes (ascending) and weights on [-1, 1], float64."""
    n = int(n)
    if n < 2:
        raise ValueError("clenshaw_curtis needs at least 2 points")
    m = n - 1
    ang = np.pi * np.arange(n, dtype=np.float64) / m
    k = np.arange(1, m // 2 + 1, dtype=np.float64)
    b = np.where(2 * k == m, 1.0, 2.0) / (4.0 * k * k - 1.0)
    series = (np.cos(2.0 * np.outer(ang, k)) * b).sum(axis=1)
    c = np.full(n, 2.0)
    c[0] = c[-1] = 1.0
    return -np.cos(ang), c / m * (1.0 - series)


def grid_1d(n_param: int, action_max: float) -> Tuple[np.ndarray, np.ndarray]:
    """action_dim == 1 grid (forwardkl_network.py:60-71): the N_param-point rule without its two end
    points (the integrand is singular there), nodes scaled by ``action_max``; both float32.
    Returns (actions [N_param-2, 1], weights [N_param-2])."""
    x, w = clenshaw_curtis(n_param)
    acts = x[1:-1].astype(np.float32)[:, None] * np.float32(action_max)
    return acts.astype(np.float32), w[1:-1].astype(np.float32)


def grid_smolyak(l_param: int, action_dim: int, action_max: float) -> Tuple[np.ndarray, np.ndarray]:
    """action_dim > 1 grid (forwardkl_network.py:73-102): Smolyak combination of the nested 1-D rules with
    1, 3, 5, 9, ... points (levels 0..l-1; every level >= 1 drops its end points, level 0 is the midpoint
    rule with weight 2).  Level multi-indices k with l <= |k| + d <= l + d - 1 contribute with the
    coefficient (-1)^(l + d - |k| - d + 1) * C(d-1, |k| + d - l), in ``itertools.product`` order, points
    inside a level in ``itertools.product`` order as well (repeated nodes are NOT merged, as in the
    reference).  Returns (actions [N, d] float32 scaled by action_max, weights [N] float32)."""
    l, d = int(l_param), int(action_dim)
    pts = [np.array([0.0])]
    wts = [np.array([2.0])]
    for i in range(1, l):
        x, w = clenshaw_curtis(2 ** i + 1)
        pts.append(x[1:-1])
        wts.append(w[1:-1])
    acts, weights = [], []
    for k in itertools.product(range(l), repeat=d):
        tot = sum(k) + d
        if tot < l or tot > l + d - 1:
            continue
        coeff = (-1) ** (l + d - sum(k) - d + 1) * comb(d - 1, sum(k) + d - l)
        for j in itertools.product(*[range(len(pts[ki])) for ki in k]):
            acts.append([np.float32(pts[k[i]][j[i]]) for i in range(d)])
            weights.append(coeff * np.prod([wts[k[i]][j[i]] for i in range(d)]))
    acts = np.asarray(acts, np.float32).reshape(-1, d) * np.float32(action_max)
    return acts.astype(np.float32), np.asarray(weights, np.float64).astype(np.float32)


def integration_grid(action_dim: int, action_max: float, n_param: int = 64, l_param: int = 6):
    """The grid a ForwardKL/ReverseKL network builds in ``__init__`` for its config."""
    if int(action_dim) == 1:
        return grid_1d(n_param, action_max)
    return grid_smolyak(l_param, action_dim, action_max)
