"""Shared test helpers: seeded instances and check lists (numpy RNG; Julia's RNG stream is
not reproducible without Julia, so instances are inputs to both the oracle and the GPU)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

POINT2D, POINT3D, ARM22 = O.POINT2D, O.POINT3D, O.ARM22

# relative offsets from the contact threshold used by the adversarial generators: rounding-level
# ones (the reference arithmetic itself decides) and ones around the 1e-6 band in which the
# kernels' division-free quick test hands over to the reference arithmetic (geom.cuh)
EPS = [0.0, 1e-16, -1e-16, 1e-15, -1e-15, 1e-13, -1e-13, 1e-10, -1e-10, 2e-7, -2e-7, 4.9e-7,
       -4.9e-7, 5.1e-7, -5.1e-7, 1e-6, -1e-6, 3e-6, -3e-6, 1e-4, -1e-4]


def make_point_instance(model, N, M, seed, rad=(0.015625, 0.0234375), rad_obs=(0.025, 0.04)):
    """Instance in the style of gen_random_instance (utils.jl:204-245) with point2d_many's
    parameter ranges (scripts/config/eval/point2d_many.yaml:12-19) by default."""
    d = 3 if model == POINT3D else 2
    rng = np.random.default_rng(seed)
    obs = np.concatenate([rng.random((M, d)), rng.uniform(*rad_obs, size=(M, 1))], axis=1) \
        if M else np.zeros((0, d + 1))
    rads = rng.uniform(*rad, size=N)
    return rads, obs


def make_arm_instance(N, M, seed, link=(0.1, 0.15), rad_obs=(0.05, 0.1)):
    rng = np.random.default_rng(seed)
    obs = np.concatenate([rng.random((M, 2)), rng.uniform(*rad_obs, size=(M, 1))], axis=1) \
        if M else np.zeros((0, 3))
    rads = rng.uniform(*link, size=N)
    base = rng.uniform(0.3, 0.7, size=(N, 2))
    return rads, obs, base


def random_edges(rng, n, d, max_len=0.1):
    """PRM-like edges: q_from uniform, q_to within max_len."""
    qf = rng.random((n, d))
    step = rng.normal(size=(n, d))
    step *= (rng.random((n, 1)) * max_len) / np.linalg.norm(step, axis=1, keepdims=True)
    qt = qf + step
    return qf, qt


def near_threshold_connect(rng, n, rads, obs, agent):
    """Adversarial edges that graze an obstacle: the segment passes the obstacle centre at
    distance (o.r + rads[i]) * (1 + tiny)."""
    d = obs.shape[1] - 1
    m = rng.integers(0, obs.shape[0], n)
    c = obs[m, :d]
    thr = obs[m, d] + rads[agent]
    dirn = rng.normal(size=(n, d))
    dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
    # a normal direction
    nrm = rng.normal(size=(n, d))
    nrm -= (nrm * dirn).sum(1, keepdims=True) * dirn
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    eps = rng.choice(EPS, n)
    foot = c + nrm * (thr * (1 + eps))[:, None]
    la, lb = rng.uniform(-0.08, 0.08, (2, n))
    return foot + dirn * la[:, None], foot + dirn * lb[:, None]


def near_threshold_pairs(rng, n, rads, ai, aj, d):
    """Two segments whose closest approach is (rads[i]+rads[j]) * (1 + tiny)."""
    thr = rads[ai] + rads[aj]
    p = rng.uniform(0.2, 0.8, (n, d))
    u = rng.normal(size=(n, d))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = rng.normal(size=(n, d))
    v -= (v * u).sum(1, keepdims=True) * u
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    eps = rng.choice(EPS, n)
    off = v * (thr * (1 + eps))[:, None]
    l1, l2, l3, l4 = rng.uniform(-0.05, 0.05, (4, n))
    w = rng.normal(size=(n, d))
    w -= (w * v).sum(1, keepdims=True) * v   # second segment direction, normal to the offset
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    return (p + u * l1[:, None], p + u * l2[:, None],
            p + off + w * l3[:, None], p + off + w * l4[:, None])


LINE2D, SNAKE2D, ARM33, CAPSULE3D = O.LINE2D, O.SNAKE2D, O.ARM33, O.CAPSULE3D
CHAIN_MODELS = (LINE2D, SNAKE2D, ARM33, CAPSULE3D)


def make_chain_instance(model, N, M, seed):
    """Instances in the spirit of the reference's eval configs (scripts/config/eval/*.yaml) with
    bodies small enough that N agents fit: returns dict(rads, obs, base, aux)."""
    rng = np.random.default_rng(seed)
    wd = O._W[model]
    obs = np.concatenate([rng.random((M, wd)), rng.uniform(0.05, 0.1, (M, 1))], axis=1) if M \
        else np.zeros((0, wd + 1))
    base = aux = None
    if model == LINE2D:
        rads = rng.uniform(0.06, 0.12, N)
    elif model == SNAKE2D:
        rads = rng.uniform(0.03, 0.05, N)
    elif model == ARM33:
        rads = rng.uniform(0.08, 0.12, N)
        base = rng.uniform(0.35, 0.65, (N, 3))
    else:
        rads = rng.uniform(0.03, 0.05, N)
        aux = rng.uniform(0.08, 0.15, N)
    return dict(rads=rads, obs=obs, base=base, aux=aux)


def chain_states(rng, model, n, near=None, spread=0.15):
    """n random states of a chain model (positions in [0.2, 0.8], angles in [0, 2pi)); with
    `near` (an array of states) perturbations of those instead."""
    sd = O._D[model]
    npos = {LINE2D: 2, SNAKE2D: 2, ARM33: 0, CAPSULE3D: 3}[model]
    if near is not None:
        q = near + rng.normal(size=near.shape) * spread
        return q
    q = rng.uniform(0, 2 * np.pi, (n, sd))
    q[:, :npos] = rng.uniform(0.2, 0.8, (n, npos))
    return q
