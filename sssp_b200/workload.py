"""Synthetic workloads of the reference's experiment configs (inputs only -- no checking).

configs[1] of BASELINE.json: StatePoint2D scalability instance of
scripts/config/eval/point2d_many.yaml (N agents, radii U[1/64, 3/128], 10 obstacles with
radii U[0.025, 0.04]; PRMs(num_vertices=500, rad=0.1)), plus the batched connect + collide
check lists of SURVEY 8(d) config 2/5.  Julia's RNG stream is not reproducible, so instances
are drawn from numpy's PCG64 with fixed seeds and handed to every implementation alike.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

POINT2D = 0


@dataclass
class Instance:
    model: int
    rads: np.ndarray       # [N]
    obstacles: np.ndarray  # [M][d+1]
    q_init: np.ndarray     # [N][d]
    q_goal: np.ndarray     # [N][d]
    nv: int                # PRM num_vertices
    rad: float             # PRM connection radius
    seed: int


def point2d_many(N: int = 50, seed: int = 0, nv: int = 500, rad: float = 0.1) -> Instance:
    """point2d_many.yaml:12-29 (instance + PP's roadmap parameters).  Starts / goals are drawn
    in the open field away from the walls; start/goal validity (utils.jl:160-201) is not part
    of the measured path, the roadmap kernels treat them as given vertices 0 and 1."""
    rng = np.random.default_rng(seed)
    M = 10
    obstacles = np.concatenate([rng.random((M, 2)), rng.uniform(0.025, 0.04, (M, 1))], axis=1)
    rads = rng.uniform(0.015625, 0.0234375, N)
    q_init = rng.uniform(0.05, 0.95, (N, 2))
    q_goal = rng.uniform(0.05, 0.95, (N, 2))
    return Instance(POINT2D, rads, obstacles, q_init, q_goal, nv, rad, seed)


def point3d_n20(N: int = 20, seed: int = 0, nv: int = 200, rad: float = 0.3) -> Instance:
    """BASELINE configs[2]: StatePoint3D, N = 20, sphere obstacles.  point3d.yaml:11-20 draws radii
    0.1-0.15 for at most 6 agents; 20 such spheres do not fit the unit cube, so the
    agents get rad 0.05 and the 10 obstacles radii U[0.05, 0.1].  PRM: num_vertices 200, rad 0.3
    (point3d.yaml:26-30)."""
    rng = np.random.default_rng(seed)
    M = 10
    obstacles = np.concatenate([rng.random((M, 3)), rng.uniform(0.05, 0.1, (M, 1))], axis=1)
    rads = np.full(N, 0.05)
    q_init = rng.uniform(0.1, 0.9, (N, 3))
    q_goal = rng.uniform(0.1, 0.9, (N, 3))
    return Instance(1, rads, obstacles, q_init, q_goal, nv, rad, seed)


def connect_list(inst: Instance, n: int, seed: int = 1234, max_len: float = 0.1):
    """n explicit connect(q_from, q_to, i) checks: endpoints U[0,1]^2, edge length <= max_len
    (PRM-like), agent ids U{0..N-1}  (SURVEY 8d, config 5)."""
    rng = np.random.default_rng(seed)
    N = inst.rads.size
    agent = rng.integers(0, N, n, dtype=np.int32)
    qf = rng.random((n, 2))
    ang = rng.random(n) * (2 * np.pi)
    ln = rng.random(n) * max_len
    qt = qf + np.stack([np.cos(ang), np.sin(ang)], axis=1) * ln[:, None]
    return agent, qf, qt


def collide_list(inst: Instance, row_ptr: np.ndarray, col: np.ndarray, n: int, seed: int = 4321):
    """n collide(6-arity) checks between random roadmap edges of random agent pairs i != j,
    referenced by (agent, vertex) ids into the per-agent roadmaps (packed CSR over all N*nv
    rows).  Rows without out-edges fall back to a stay edge v -> v."""
    rng = np.random.default_rng(seed)
    N, nv = inst.rads.size, inst.nv
    ai = rng.integers(0, N, n, dtype=np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    row_ptr = row_ptr.astype(np.int64)

    def edges(a):
        vf = rng.integers(0, nv, n, dtype=np.int32)
        r = a.astype(np.int64) * nv + vf
        deg = row_ptr[r + 1] - row_ptr[r]
        pick = (rng.random(n) * np.maximum(deg, 1)).astype(np.int64)
        vt = np.where(deg > 0, col[np.minimum(row_ptr[r] + pick, col.size - 1)], vf)
        return vf, vt.astype(np.int32)

    vif, vit = edges(ai)
    vjf, vjt = edges(aj)
    return ai, aj, vif, vit, vjf, vjt


def random_walk_paths(inst: Instance, row_ptr: np.ndarray, col: np.ndarray, T: int, seed: int):
    """Synthetic planned paths: for every agent a T-step random walk on its own roadmap from
    vertex 0 (its start, prm.jl:232); an agent without out-edges waits.  Returns the vertex
    table [N][T+1] and the path-edge columns in t-major order (column t*N + j is agent j's
    move at step t), i.e. what PP's invalid() iterates over (pp.jl:179-189)."""
    rng = np.random.default_rng(seed)
    N, nv = inst.rads.size, inst.nv
    row_ptr = row_ptr.astype(np.int64)
    paths = np.zeros((N, T + 1), np.int32)
    v = np.zeros(N, np.int64)
    base = np.arange(N, dtype=np.int64) * nv
    for t in range(T):
        r = base + v
        deg = row_ptr[r + 1] - row_ptr[r]
        pick = (rng.random(N) * np.maximum(deg, 1)).astype(np.int64)
        nxt = np.where(deg > 0, col[np.minimum(row_ptr[r] + pick, col.size - 1)], v)
        v = nxt.astype(np.int64)
        paths[:, t + 1] = v
    col_agent = np.tile(np.arange(N, dtype=np.int32), T)
    col_vfrom = np.ascontiguousarray(paths[:, :-1].T).reshape(-1).astype(np.int32)
    col_vto = np.ascontiguousarray(paths[:, 1:].T).reshape(-1).astype(np.int32)
    return paths, col_agent, col_vfrom, col_vto


def gate_pass_count(nodes: np.ndarray, rad: float) -> int:
    """Number of ordered vertex pairs (j != k) within `rad`: the connect(q_j, q_k) calls the
    reference's PRM! loop actually issues (prm.jl:207)."""
    tot = 0
    for a in range(nodes.shape[0]):
        d = nodes[a][:, None, :] - nodes[a][None, :, :]
        tot += int((np.sqrt((d * d).sum(-1)) <= rad).sum()) - nodes.shape[1]
    return tot
