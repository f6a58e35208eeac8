"""Host-side mirror of the reference's closure API on top of the C ABI.

Reference surface (Kei18/sssp, paths under /root/reference):
  gen_connect(q, obstacles, ins_params...; step_dist, max_dist, safety_dist) -> f
      f(q, i), f(q_from, q_to, i)                 point2d.jl:42-69, point3d.jl:47-80, arm22.jl:29-88
  gen_collide(q, ins_params...; step_dist, safety_dist) -> f
      f(qi_from, qi_to, qj_from, qj_to, i, j), f(q_i, q_j, i, j),
      f(Q_from, Q_to), f(Q, q_to, i)              point.jl:5-36, arm22.jl:90-212
  gen_check_goal(config_goal; goal_rad, goal_rads) -> f     utils.jl:117-138
  PRMs(config_init, config_goal, connect, num_vertices; rad, rads)   prm.jl:251-276

Differences from Julia, all mechanical: agent indices and vertex ids are 0-based here;
methods that Julia selects by arity/type are selected the same way (argument count and
whether the first argument is a configuration); every closure additionally exposes
batched entry points (`.points`, `.edges`, `.pairs`, ...) which are the new API surface
the GPU path adds.  All verdicts come from the CUDA library -- no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import sys
from dataclasses import dataclass, field
from typing import List, NamedTuple, Optional, Sequence

import numpy as np

from . import lib as L

STEP_DIST = 0.01         # src/MRMP.jl:65
SAFETY_DIST_LINE = 0.01  # src/MRMP.jl:66


class StatePoint2D(NamedTuple):  # src/models/point2d.jl:3-6
    x: float
    y: float


class StatePoint3D(NamedTuple):  # src/models/point3d.jl:3-7
    x: float
    y: float
    z: float


class StateArm22(NamedTuple):  # src/models/arm22.jl:3-6
    theta1: float
    theta2: float


class StateLine2D(NamedTuple):  # src/models/line2d.jl:3-7
    x: float
    y: float
    theta: float


class StateSnake2D(NamedTuple):  # src/models/snake2d.jl:5-12
    x: float
    y: float
    theta1: float
    theta2: float
    theta3: float
    theta4: float


class StateArm33(NamedTuple):  # src/models/arm33.jl:3-10
    theta1: float
    phi1: float
    theta2: float
    phi2: float
    theta3: float
    phi3: float


class StateCapsule3D(NamedTuple):  # src/models/capsule3d.jl:5-12
    x: float
    y: float
    z: float
    phi: float
    theta: float
    psi: float


STEP_DIST_SNAKE2D = 0.05    # src/models/snake2d.jl:3
STEP_DIST_CAPSULE3D = 0.05  # src/models/capsule3d.jl:3


class CircleObstacle2D(NamedTuple):  # src/obstacle.jl:7-11
    x: float
    y: float
    r: float


class CircleObstacle3D(NamedTuple):  # src/obstacle.jl:14-19
    x: float
    y: float
    z: float
    r: float


@dataclass
class Node:  # src/MRMP.jl:73-77
    q: tuple
    id: int
    neighbors: List[int] = field(default_factory=list)


_MODEL_OF = {StatePoint2D: L.MODEL_POINT2D, StatePoint3D: L.MODEL_POINT3D,
             StateArm22: L.MODEL_ARM22, StateLine2D: L.MODEL_LINE2D, StateSnake2D: L.MODEL_SNAKE2D,
             StateArm33: L.MODEL_ARM33, StateCapsule3D: L.MODEL_CAPSULE3D}
_STATE_OF = {m: t for t, m in _MODEL_OF.items()}
_DEFAULT_STEP = {L.MODEL_SNAKE2D: STEP_DIST_SNAKE2D, L.MODEL_CAPSULE3D: STEP_DIST_CAPSULE3D}


def _model_of(q) -> int:
    for t, m in _MODEL_OF.items():
        if isinstance(q, t) or q is t:
            return m
    raise TypeError(f"unsupported state type {type(q).__name__}")


def _states(qs, d: int) -> np.ndarray:
    """Vector{State} / Vector{Node} / ndarray -> contiguous float64 [n][d]."""
    if isinstance(qs, np.ndarray):
        return L.f64(qs).reshape(-1, d)
    rows = [(q.q if isinstance(q, Node) else q) for q in qs]
    return L.f64(rows).reshape(-1, d)


class Context:
    """One problem instance on one GPU (wraps mrmp_ctx)."""

    def __init__(self, model: int, rads, obstacles=(), base_pos=None, step_dist=None,
                 safety_dist=SAFETY_DIST_LINE, max_dist: Optional[float] = None, device: int = 0,
                 aux=None):
        self.model = int(model)
        self.d = L.STATE_DIM[self.model]
        self.wd = L.WORK_DIM[self.model]
        self.rads = L.f64(rads).ravel()
        self.N = int(self.rads.size)
        obs = L.f64([tuple(o) for o in obstacles] if not isinstance(obstacles, np.ndarray)
                    else obstacles).reshape(-1, self.wd + 1)
        self.obstacles = obs
        base = None if base_pos is None else L.f64(base_pos).reshape(self.N, self.wd)
        self.base_pos = base
        self.aux = None if aux is None else L.f64(aux).ravel()
        if step_dist is None:
            step_dist = _DEFAULT_STEP.get(self.model, STEP_DIST)
        self.device = int(device)
        self.step_dist, self.safety_dist = float(step_dist), float(safety_dist)
        self.max_dist = None if max_dist is None else float(max_dist)
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.mrmp_ctx_create(
            C.byref(h), self.model, self.N, L.ptr(self.rads, L._dp),
            L.ptr(base, L._dp) if base is not None else None,
            L.ptr(self.aux, L._dp) if self.aux is not None else None, obs.shape[0],
            L.ptr(obs, L._dp) if obs.size else None, float(step_dist), float(safety_dist),
            math.nan if max_dist is None else float(max_dist), self.device))
        self._h = h
        # Objects created on this context (Roadmaps, SsspRoadmaps, peer.Exchange) use it in their
        # own destroy call, and the cyclic garbage collector finalises in no particular order:
        # the context is only destroyed once the last of them is gone.
        self._dependents = 0
        self._close_wanted = False

    def _adopt(self) -> None:
        self._dependents += 1

    def _release(self) -> None:
        self._dependents -= 1
        if self._dependents <= 0 and self._close_wanted:
            self.close()

    def close(self):
        if getattr(self, "_dependents", 0) > 0:
            self._close_wanted = True      # deferred until the dependents are destroyed
            return
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mrmp_ctx_destroy(h)

    def __del__(self):
        # at interpreter shutdown the CUDA runtime may already be gone: leave the handle to the
        # OS instead of calling into the library from a finaliser
        if sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def stats(self, reset: bool = False) -> np.ndarray:
        out = np.zeros(4, np.int64)
        L.check(self._lib.mrmp_ctx_stats(self._h, L.ptr(out, L._vp), int(reset)))
        return out

    def stats8(self, reset: bool = False) -> np.ndarray:
        """the four counters of stats(), then [4] arm22 link-pair midpoint bounds evaluated"""
        out = np.zeros(8, np.int64)
        L.check(self._lib.mrmp_ctx_stats8(self._h, L.ptr(out, L._vp), int(reset)))
        return out

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run this ctx's kernels on a caller-owned CUDA stream (e.g. torch's current); None
        restores the ctx's own stream.  The handle of the default stream is 0 (what
        ``torch.cuda.current_stream().cuda_stream`` returns outside a stream context), which the
        library would read as "own stream": it is passed on as cudaStreamLegacy (0x1), so that the
        ctx's launches and the caller's work on stream 0 are ordered on ONE stream."""
        if cuda_stream is not None and int(cuda_stream) == 0:
            cuda_stream = 1   # cudaStreamLegacy
        L.check(self._lib.mrmp_ctx_set_stream(self._h, cuda_stream))

    # ---- batched host-buffer entry points (numpy in / numpy out) ---------------------
    def connect_points(self, agent, q) -> np.ndarray:
        agent, q = L.i32(agent).ravel(), _states(q, self.d)
        out = np.empty(agent.size, np.uint8)
        L.check(self._lib.mrmp_connect_points(self._h, agent.size, L.ptr(agent, L._ip),
                                              L.ptr(q, L._dp), L.ptr(out, L._bp)))
        return out

    def connect_edges(self, agent, q_from, q_to) -> np.ndarray:
        agent = L.i32(agent).ravel()
        qf, qt = _states(q_from, self.d), _states(q_to, self.d)
        out = np.empty(agent.size, np.uint8)
        L.check(self._lib.mrmp_connect_edges(self._h, agent.size, L.ptr(agent, L._ip),
                                             L.ptr(qf, L._dp), L.ptr(qt, L._dp),
                                             L.ptr(out, L._bp)))
        return out

    def collide_pairs(self, i, j, qi_from, qi_to, qj_from, qj_to) -> np.ndarray:
        i, j = L.i32(i).ravel(), L.i32(j).ravel()
        a, b, c, d = (_states(x, self.d) for x in (qi_from, qi_to, qj_from, qj_to))
        out = np.empty(i.size, np.uint8)
        L.check(self._lib.mrmp_collide_pairs(self._h, i.size, L.ptr(i, L._ip), L.ptr(j, L._ip),
                                             L.ptr(a, L._dp), L.ptr(b, L._dp), L.ptr(c, L._dp),
                                             L.ptr(d, L._dp), L.ptr(out, L._bp)))
        return out

    def collide_static_pairs(self, i, j, qi, qj) -> np.ndarray:
        i, j = L.i32(i).ravel(), L.i32(j).ravel()
        a, b = _states(qi, self.d), _states(qj, self.d)
        out = np.empty(i.size, np.uint8)
        L.check(self._lib.mrmp_collide_static_pairs(self._h, i.size, L.ptr(i, L._ip),
                                                    L.ptr(j, L._ip), L.ptr(a, L._dp),
                                                    L.ptr(b, L._dp), L.ptr(out, L._bp)))
        return out

    def collide_configs(self, Q_from, Q_to):
        """n_cfg x N states each -> (verdict[n_cfg], first_pair[n_cfg] = i*N+j or -1)"""
        a, b = L.f64(Q_from).reshape(-1, self.N, self.d), L.f64(Q_to).reshape(-1, self.N, self.d)
        n = a.shape[0]
        out = np.empty(n, np.uint8)
        fp = np.empty(n, np.int64)
        L.check(self._lib.mrmp_collide_configs(self._h, n, L.ptr(a, L._dp), L.ptr(b, L._dp),
                                               L.ptr(out, L._bp), L.ptr(fp, L._lp)))
        return out, fp

    def collide_config_moves(self, agent_i: int, Q, q_to) -> np.ndarray:
        Q, q_to = _states(Q, self.d), _states(q_to, self.d)
        out = np.empty(q_to.shape[0], np.uint8)
        L.check(self._lib.mrmp_collide_config_moves(self._h, int(agent_i), L.ptr(Q, L._dp),
                                                    q_to.shape[0], L.ptr(q_to, L._dp),
                                                    L.ptr(out, L._bp)))
        return out

    def collide_cross(self, row_agent, row_from, row_to, col_agent, col_from, col_to,
                      group_size: int = 0) -> np.ndarray:
        """Every row edge x every column edge -> bit matrix uint32 [n_rows][words_per_row]
        (group_size > 0: OR-reduced over consecutive column groups).  Same-agent pairs are 0."""
        ra, ca = L.i32(row_agent).ravel(), L.i32(col_agent).ravel()
        rf, rt = _states(row_from, self.d), _states(row_to, self.d)
        cf, ct = _states(col_from, self.d), _states(col_to, self.d)
        bits = -(-ca.size // group_size) if group_size > 0 else ca.size
        wpr = (bits + 31) // 32
        out = np.zeros((ra.size, max(wpr, 0)), np.uint32)
        L.check(self._lib.mrmp_collide_cross(self._h, ra.size, L.ptr(ra, L._ip), L.ptr(rf, L._dp),
                                             L.ptr(rt, L._dp), ca.size, L.ptr(ca, L._ip),
                                             L.ptr(cf, L._dp), L.ptr(ct, L._dp), int(group_size),
                                             L.ptr(out, L._vp)))
        return out

    def collide_cross_reduce(self, row_agent, row_from, row_to, col_agent, col_from, col_to, mode,
                             group_size: int = 0, col_group=None, n_groups=None, row_key=None,
                             col_key=None, cbs: bool = False) -> np.ndarray:
        """Reductions of the cross product -> int32 [n_rows][n_groups].
        mode "count": colliding columns per (row, group); with cbs=True the pair test of CBS's soft
        conflict count (cbs.jl:336-347: vertex conflict OR edge conflict, column agent first).
        mode "first": smallest colliding column index per (row, group), -1 if none; row_key /
        col_key restrict to col_key > row_key (TPG type-2 dependencies, smoother.jl:78-106)."""
        ra, ca = L.i32(row_agent).ravel(), L.i32(col_agent).ravel()
        rf, rt = _states(row_from, self.d), _states(row_to, self.d)
        cf, ct = _states(col_from, self.d), _states(col_to, self.d)
        cg = L.i32(col_group).ravel() if col_group is not None else None
        rk = L.i32(row_key).ravel() if row_key is not None else None
        ck = L.i32(col_key).ravel() if col_key is not None else None
        if n_groups is None:
            n_groups = (int(cg.max()) + 1 if cg.size else 1) if cg is not None else max(1, -(-ca.size // group_size))
        out = np.zeros((ra.size, n_groups), np.int32)
        vp = lambda a: a.ctypes.data_as(L._vp) if a is not None else None
        L.check(self._lib.mrmp_collide_cross_reduce(
            self._h, ra.size, vp(ra), vp(rf), vp(rt), vp(rk), ca.size, vp(ca), vp(cf), vp(ct), vp(cg),
            vp(ck), int(group_size), int(n_groups), {"count": L.CROSS_COUNT, "first": L.CROSS_FIRST}[mode],
            L.CROSS_FLAG_CBS if cbs else 0, vp(out)))
        return out

    def sample_states(self, agent: int, seed: int, t0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.d))
        L.check(self._lib.mrmp_sample_states(self._h, int(agent), int(seed), int(t0), int(n),
                                             L.ptr(out, L._dp)))
        return out

    def sample_free(self, agent: int, seed: int, n_wanted: int):
        out = np.empty((n_wanted, self.d))
        tried = C.c_int64(-1)
        L.check(self._lib.mrmp_sample_free(self._h, int(agent), int(seed), int(n_wanted),
                                           L.ptr(out, L._dp), C.byref(tried)))
        return out, int(tried.value)

    # ---- model helpers used by the host search loops ----------------------------------
    def state_dist(self, q1, q2) -> np.ndarray:
        """dist(q1[k], q2[k]) of the model (point2d.jl:12-14, point3d.jl:13-15, arm22.jl:15-20)"""
        a, b = _states(q1, self.d), _states(q2, self.d)
        out = np.empty(a.shape[0])
        L.check(self._lib.mrmp_state_dist(self._h, a.shape[0], L.ptr(a, L._dp), L.ptr(b, L._dp),
                                          L.ptr(out, L._dp)))
        return out

    def mid_status(self, p, q) -> np.ndarray:
        """get_mid_status(p[k], q[k]) (point2d.jl:8-10, point3d.jl:9-11, arm22.jl:8-13)"""
        a, b = _states(p, self.d), _states(q, self.d)
        out = np.empty_like(a)
        L.check(self._lib.mrmp_mid_status(self._h, a.shape[0], L.ptr(a, L._dp), L.ptr(b, L._dp),
                                          L.ptr(out, L._dp)))
        return out

    def nearest(self, V, q, reversed: bool = False):
        """argmin_v dist(V[v], q) (reversed: dist(q, V[v])) -> (index, distance)"""
        V, q = _states(V, self.d), _states([q] if not isinstance(q, np.ndarray) else q, self.d)
        idx, dist = C.c_int32(-1), C.c_double(0.0)
        L.check(self._lib.mrmp_nearest(self._h, V.shape[0], L.ptr(V, L._dp), L.ptr(q, L._dp),
                                       int(reversed), C.byref(idx), C.byref(dist)))
        return int(idx.value), float(dist.value)

    def rrt_extend(self, agent: int, V, q_rand, from_start: bool, steering_depth: int,
                   epsilon: Optional[float]):
        """extend! of RRT-Connect (sssp.jl:404-440) in one call -> (flag, q_new, index of q_near);
        flag in {"trapped", "advanced", "reached"}."""
        V = _states(V, self.d)
        q = L.f64(q_rand).ravel()
        flag, near = C.c_int32(-1), C.c_int32(-1)
        out = np.empty(self.d)
        L.check(self._lib.mrmp_rrt_extend(self._h, int(agent), V.shape[0], L.ptr(V, L._dp),
                                          L.ptr(q, L._dp), int(bool(from_start)), int(steering_depth),
                                          math.nan if epsilon is None else float(epsilon),
                                          C.byref(flag), C.byref(near), L.ptr(out, L._dp)))
        return ("trapped", "advanced", "reached")[flag.value], out, int(near.value)

    def first_conflict(self, configs):
        """get_constraints' scan (cbs.jl:374-412) over configs [T][N][d]: None when conflict
        free, else (t, kind, i, j) with kind 'vertex' / 'edge' of the first conflict."""
        cfg = L.f64(configs).reshape(-1, self.N, self.d)
        t, kind, i, j = C.c_int64(-1), C.c_int32(0), C.c_int32(-1), C.c_int32(-1)
        L.check(self._lib.mrmp_first_conflict(self._h, cfg.shape[0], L.ptr(cfg, L._dp), C.byref(t),
                                              C.byref(kind), C.byref(i), C.byref(j)))
        if t.value < 0:
            return None
        return int(t.value), "vertex" if kind.value == 1 else "edge", int(i.value), int(j.value)

    def distance_tables(self, graphs, goal: int = 1, stay_penalty: float = 0.0):
        """get_distance_table for several roadmaps in one launch (solvers/utils.jl:146-236).
        graphs: list of (nodes [nv][d], neighbour lists) ; returns a list of float64 tables.
        stay_penalty: the edge cost is gen_g_func(stay_penalty)'s (solvers/utils.jl:34-36)."""
        v_off = np.zeros(len(graphs) + 1, np.int32)
        nodes, rp, ci = [], [np.zeros(1, np.int64)], []
        for g, (q, nbrs) in enumerate(graphs):
            q = _states(q, self.d)
            v_off[g + 1] = v_off[g] + q.shape[0]
            nodes.append(q)
            deg = np.fromiter((len(x) for x in nbrs), np.int64, count=len(nbrs))
            rp.append(rp[-1][-1] + np.cumsum(deg))
            ci.append(np.fromiter((u for x in nbrs for u in x), np.int32, count=int(deg.sum())))
        nodes = np.ascontiguousarray(np.concatenate(nodes))
        row_ptr = np.concatenate(rp).astype(np.int32)
        col = np.ascontiguousarray(np.concatenate(ci)) if ci else np.zeros(0, np.int32)
        goals = np.full(len(graphs), goal, np.int32)
        out = np.empty(nodes.shape[0])
        L.check(self._lib.mrmp_distance_tables_g(self._h, len(graphs), L.ptr(v_off, L._ip),
                                                 L.ptr(nodes, L._dp), L.ptr(row_ptr, L._ip),
                                                 L.ptr(col, L._ip) if col.size else None,
                                                 L.ptr(goals, L._ip), float(stay_penalty),
                                                 L.ptr(out, L._dp)))
        return [out[v_off[g]: v_off[g + 1]].copy() for g in range(len(graphs))]

    def prm_build(self, nodes, rad, agent0: int = 0):
        """nodes [n][nv][d], rad scalar/array/None -> (row_ptr [n*nv+1], col_idx [nnz])"""
        nodes = L.f64(nodes)
        n, nv = nodes.shape[0], nodes.shape[1]
        radv = np.broadcast_to(L.f64(math.nan if rad is None else rad), (n,)).copy()
        row_ptr = np.empty(n * nv + 1, np.int32)
        cap = max(1, n * nv * 64)
        nnz = C.c_int64(0)
        while True:
            col = np.empty(cap, np.int32)
            rc = self._lib.mrmp_prm_build(self._h, int(agent0), n, nv, L.ptr(radv, L._dp),
                                          L.ptr(nodes, L._dp), L.ptr(row_ptr, L._ip),
                                          L.ptr(col, L._ip), cap, C.byref(nnz))
            if rc == L.ERR_CAPACITY and nnz.value > cap:
                cap = int(nnz.value)
                continue
            L.check(rc)
            return row_ptr, col[: nnz.value].copy()


def collide_params_agree(a: "Context", b: "Context") -> bool:
    """True when two contexts give the same collide verdicts: same model, radii, base positions,
    capsule axes, step_dist and safety_dist (the parameters gen_collide captures: point.jl:5-8,
    arm22.jl:90-99, ...).  Obstacles and max_dist only matter to connect."""
    def same(x, y):
        return (x is None and y is None) or (x is not None and y is not None and np.array_equal(x, y))
    return (a.model == b.model and a.N == b.N and np.array_equal(a.rads, b.rads)
            and same(a.base_pos, b.base_pos) and same(a.aux, b.aux)
            and a.step_dist == b.step_dist and a.safety_dist == b.safety_dist)


class Roadmaps:
    """Device-resident roadmaps of agents [agent0, agent0+n) (wraps mrmp_roadmaps)."""

    def __init__(self, ctx: Context, agent0: int, n_agents: int, nv_cap: int):
        self.ctx, self.agent0, self.n, self.nv_cap = ctx, int(agent0), int(n_agents), int(nv_cap)
        self._lib = ctx._lib
        h = C.c_void_p()
        L.check(self._lib.mrmp_roadmaps_create(ctx._h, self.agent0, self.n, self.nv_cap,
                                               C.byref(h)))
        self._h = h
        self.nv = 0
        ctx._adopt()

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mrmp_roadmaps_destroy(h)
            self.ctx._release()

    def __del__(self):
        # at interpreter shutdown the CUDA runtime may already be gone: leave the handle to the
        # OS instead of calling into the library from a finaliser
        if sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def set_nodes(self, nodes) -> None:
        nodes = L.f64(nodes).reshape(self.n, -1, self.ctx.d)
        self.nv = nodes.shape[1]
        L.check(self._lib.mrmp_roadmaps_set_nodes(self._h, self.nv, L.ptr(nodes, L._dp)))

    def bind_nodes(self, nv: int, nodes_dev_ptr: int | None) -> None:
        """Use caller-owned device memory [n][nv][d] as the node table (NCCL buffer slice)."""
        L.check(self._lib.mrmp_roadmaps_bind_nodes(self._h, int(nv), nodes_dev_ptr))
        self.nv = int(nv)

    def sample(self, nv: int, seed: int, q_init, q_goal, tries: bool = True):
        """Draw the vertex tables on the device. With ``tries=False`` the call does not wait for
        the device: a sampler that gave up is then reported by the next build()/nodes()."""
        qi, qg = _states(q_init, self.ctx.d), _states(q_goal, self.ctx.d)
        out = np.empty(self.n, np.int64) if tries else None
        L.check(self._lib.mrmp_roadmaps_sample(self._h, int(nv), int(seed), L.ptr(qi, L._dp),
                                               L.ptr(qg, L._dp),
                                               L.ptr(out, L._lp) if tries else None))
        self.nv = int(nv)
        return out

    def build(self, rad) -> int:
        radv = np.broadcast_to(L.f64(math.nan if rad is None else rad), (self.n,)).copy()
        L.check(self._lib.mrmp_roadmaps_build(self._h, L.ptr(radv, L._dp)))
        nnz = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_nnz(self._h, C.byref(nnz)))
        return int(nnz.value)

    def build_async(self, rad) -> None:
        """Launch the neighbour pass without waiting for the edge count (mrmp_roadmaps_build_async);
        until finish() the set may only feed edge_conflicts_async / the exchange pushes."""
        radv = np.broadcast_to(L.f64(math.nan if rad is None else rad), (self.n,)).copy()
        L.check(self._lib.mrmp_roadmaps_build_async(self._h, L.ptr(radv, L._dp)))

    def finish(self) -> int:
        """Wait for an asynchronous build / adoption, validate capacities, return nnz."""
        nnz = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_finish(self._h, C.byref(nnz)))
        return int(nnz.value)

    def used_grid(self) -> bool:
        """True when the last build() searched neighbours through the uniform-grid hash."""
        return bool(self._lib.mrmp_roadmaps_used_grid(self._h))

    def nodes(self) -> np.ndarray:
        out = np.empty((self.n, self.nv, self.ctx.d))
        L.check(self._lib.mrmp_roadmaps_get_nodes(self._h, L.ptr(out, L._dp)))
        return out

    def csr(self):
        nnz = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_nnz(self._h, C.byref(nnz)))
        row_ptr = np.empty(self.n * self.nv + 1, np.int32)
        col = np.empty(max(1, nnz.value), np.int32)
        L.check(self._lib.mrmp_roadmaps_get_csr(self._h, L.ptr(row_ptr, L._ip), L.ptr(col, L._ip),
                                                col.size, C.byref(nnz)))
        return row_ptr, col[: nnz.value]

    def device_ptrs(self):
        nv = C.c_int(0)
        nodes, rp, ci = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_device_ptrs(self._h, C.byref(nv), C.byref(nodes),
                                                    C.byref(rp), C.byref(ci), C.byref(nnz)))
        return dict(nv=nv.value, nodes=nodes.value, row_ptr=rp.value, col_idx=ci.value,
                    nnz=nnz.value)

    def collide_edges_indexed(self, i, j, vi_from, vi_to, vj_from, vj_to) -> np.ndarray:
        a = [L.i32(x).ravel() for x in (i, j, vi_from, vi_to, vj_from, vj_to)]
        out = np.empty(a[0].size, np.uint8)
        L.check(self._lib.mrmp_collide_edges_indexed(self._h, a[0].size,
                                                     *[L.ptr(x, L._ip) for x in a],
                                                     L.ptr(out, L._bp)))
        return out

    def edge_conflicts(self, col_agent, col_vfrom, col_vto, group_size: int = 0,
                       cols: "Roadmaps | None" = None) -> np.ndarray:
        """All roadmap edges (packed CSR order) x the given roadmap edges (by ids) -> bits
        uint32 [nnz][words_per_row]; the conflict table PP / CBS consult (pp.jl:179-189)."""
        ca, vf, vt = (L.i32(x).ravel() for x in (col_agent, col_vfrom, col_vto))
        nnz = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_nnz(self._h, C.byref(nnz)))
        bits = -(-ca.size // group_size) if group_size > 0 else ca.size
        wpr = (bits + 31) // 32
        out = np.zeros((nnz.value, wpr), np.uint32)
        L.check(self._lib.mrmp_roadmaps_edge_conflicts(self._h, cols._h if cols else None,
                                                       ca.size, L.ptr(ca, L._ip),
                                                       L.ptr(vf, L._ip), L.ptr(vt, L._ip),
                                                       int(group_size), L.ptr(out, L._vp),
                                                       out.size))
        return out

    def edge_conflicts_dev(self, n_cols: int, col_agent_ptr: int, col_vfrom_ptr: int,
                           col_vto_ptr: int, group_size: int, out_ptr: int, stream: int,
                           cols: "Roadmaps | None" = None) -> None:
        """The same table with device-resident column ids and a device (or peer-mapped) output
        buffer, asynchronous on `stream`: a rank of a multi-GPU job computes the rows of its agents
        against all agents' path edges (cols = the full set) and writes them straight into the
        memory of the rank that gathers the table (sssp_b200/peer.py)."""
        L.check(self._lib.mrmp_roadmaps_edge_conflicts_dev(self._h, cols._h if cols else None,
                                                           int(n_cols), col_agent_ptr,
                                                           col_vfrom_ptr, col_vto_ptr,
                                                           int(group_size), out_ptr, stream))

    def edge_conflicts_async(self, n_cols: int, col_agent_ptr: int, col_vfrom_ptr: int,
                             col_vto_ptr: int, group_size: int, out_ptr: int, out_cap_rows: int,
                             stream: int, cols: "Roadmaps | None" = None) -> None:
        """edge_conflicts_dev behind build_async: the row count stays on the device; `out` has
        room for out_cap_rows rows (checked by finish())."""
        L.check(self._lib.mrmp_roadmaps_edge_conflicts_async(self._h, cols._h if cols else None,
                                                             int(n_cols), col_agent_ptr,
                                                             col_vfrom_ptr, col_vto_ptr,
                                                             int(group_size), out_ptr,
                                                             int(out_cap_rows), stream))

    def to_nodes(self) -> List[List[Node]]:
        """Vector{Vector{Node}} as the reference's PRMs returns it (0-based ids)."""
        tabs = self.nodes()
        row_ptr, col = self.csr()
        State = _STATE_OF[self.ctx.model]
        res = []
        for a in range(self.n):
            rm = []
            for v in range(self.nv):
                r = a * self.nv + v
                rm.append(Node(State(*tabs[a, v]), v, col[row_ptr[r]: row_ptr[r + 1]].tolist()))
            res.append(rm)
        return res


class SsspRoadmaps:
    """Device-resident growing roadmaps of all agents + the fused node expansion
    (wraps mrmp_sssp; sssp.jl:181-223, 235-291)."""

    def __init__(self, ctx: Context, nv_cap: int = 256):
        self.ctx, self._lib = ctx, ctx._lib
        h = C.c_void_p()
        L.check(self._lib.mrmp_sssp_create(ctx._h, int(nv_cap), C.byref(h)))
        self._h = h
        ctx._adopt()

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mrmp_sssp_destroy(h)
            self.ctx._release()

    def __del__(self):
        # at interpreter shutdown the CUDA runtime may already be gone: leave the handle to the
        # OS instead of calling into the library from a finaliser
        if sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def set_roadmap(self, agent: int, nodes) -> None:
        nodes = _states(nodes, self.ctx.d)
        L.check(self._lib.mrmp_sssp_set_roadmap(self._h, int(agent), nodes.shape[0],
                                                L.ptr(nodes, L._dp)))

    def nv(self, agent: int) -> int:
        n = C.c_int(0)
        L.check(self._lib.mrmp_sssp_nv(self._h, int(agent), C.byref(n)))
        return int(n.value)

    def nodes(self, agent: int) -> np.ndarray:
        out = np.empty((self.nv(agent), self.ctx.d))
        L.check(self._lib.mrmp_sssp_get_nodes(self._h, int(agent), L.ptr(out, L._dp), out.shape[0]))
        return out

    def expand(self, agent: int, v_from: int, Q_ids, old_nbrs, n_samples: int, *, seed: int = 0,
               t0: int = 0, q_rand=None, steering_depth: int = 2, epsilon: Optional[float] = None,
               min_dist_thread: float = 0.1, slow_collide: bool = False):
        """One node expansion.  Returns (q_new [n_new][d], out_lists, in_lists, succ_ids,
        succ_collide): out_lists[r] / in_lists[r] are the ascending vertex ids the new vertex r
        reaches / is reached from (its own id closes in_lists[r] when the self loop passes)."""
        K, d = int(n_samples), self.ctx.d
        Q_ids, old = L.i32(Q_ids).ravel(), L.i32(old_nbrs).ravel()
        nv0 = self.nv(agent)
        words = (nv0 + K + 31) // 32
        q_new = np.empty((K, d))
        out_bits = np.zeros((K, words), np.uint32)
        in_bits = np.zeros((K, words), np.uint32)
        cap_s = old.size + K + 1
        succ, verdict = np.empty(cap_s, np.int32), np.empty(cap_s, np.uint8)
        n_new, n_succ, w = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        qr = None if q_rand is None else _states(q_rand, d)
        L.check(self._lib.mrmp_sssp_expand(
            self._h, int(agent), int(v_from), K, None if qr is None else L.ptr(qr, L._dp),
            int(seed), int(t0), int(steering_depth),
            math.nan if epsilon is None else float(epsilon), float(min_dist_thread),
            L.ptr(Q_ids, L._ip), old.size, L.ptr(old, L._ip) if old.size else None,
            int(slow_collide), C.byref(n_new), L.ptr(q_new, L._dp), C.byref(w),
            L.ptr(out_bits, L._vp), L.ptr(in_bits, L._vp), out_bits.size, C.byref(n_succ),
            L.ptr(succ, L._ip), L.ptr(verdict, L._bp)))
        n = int(n_new.value)

        def ids(row):
            return np.flatnonzero(np.unpackbits(row.view(np.uint8), bitorder="little")).tolist()

        return (q_new[:n], [ids(out_bits[r]) for r in range(n)], [ids(in_bits[r]) for r in range(n)],
                succ[: n_succ.value].copy(), verdict[: n_succ.value].copy())


# ---------------------------------------------------------------------------------------
# closure constructors
# ---------------------------------------------------------------------------------------
def _split_ins_params(model: int, ins_params):
    """ins_params of the reference's constructors -> (rads, positions, axises)."""
    if model in (L.MODEL_ARM22, L.MODEL_ARM33):   # (positions, rads)  arm22.jl:29-37, arm33.jl:42-50
        positions, rads = ins_params
        return L.f64(rads).ravel(), L.f64(positions).reshape(-1, L.WORK_DIM[model]), None
    if model == L.MODEL_CAPSULE3D:                # (rads, axises)  capsule3d.jl:61-68
        rads, axises = ins_params
        return L.f64(rads).ravel(), None, L.f64(axises).ravel()
    (rads,) = ins_params                          # (rads,)  point2d.jl:42-46, line2d.jl:22-28
    return L.f64(rads).ravel(), None, None


def gen_connect(q, obstacles, *ins_params, step_dist: Optional[float] = None,
                max_dist: Optional[float] = None, safety_dist: float = SAFETY_DIST_LINE,
                device: int = 0):
    """gen_connect(q, obstacles, ins_params...) -> f;  f(q, i) / f(q_from, q_to, i).
    step_dist defaults to the model's constant (STEP_DIST, or 0.05 for snake2d / capsule3d)."""
    model = _model_of(q)
    rads, base, aux = _split_ins_params(model, ins_params)
    ctx = Context(model, rads, obstacles, base, step_dist, safety_dist, max_dist, device, aux=aux)

    def f(*args) -> bool:
        if len(args) == 2:    # f(q, i)
            return bool(ctx.connect_points([args[1]], [args[0]])[0])
        if len(args) == 3:    # f(q_from, q_to, i)
            return bool(ctx.connect_edges([args[2]], [args[0]], [args[1]])[0])
        raise TypeError("connect takes (q, i) or (q_from, q_to, i)")

    f.ctx = ctx
    f.points = ctx.connect_points   # batched f(q, i)
    f.edges = ctx.connect_edges     # batched f(q_from, q_to, i)
    return f


def gen_collide(q, *ins_params, step_dist: Optional[float] = None,
                safety_dist: float = SAFETY_DIST_LINE, device: int = 0):
    """gen_collide(q, ins_params...) -> f with the reference's four methods."""
    model = _model_of(q)
    rads, base, aux = _split_ins_params(model, ins_params)
    ctx = Context(model, rads, (), base, step_dist, safety_dist, None, device, aux=aux)
    State = _STATE_OF[model]

    def _is_config(x) -> bool:
        return (isinstance(x, (list, tuple, np.ndarray)) and not isinstance(x, State)
                and len(x) > 0 and isinstance(x[0], (Node, State, tuple, list, np.ndarray)))

    def f(*args) -> bool:
        if len(args) == 6:    # f(qi_from, qi_to, qj_from, qj_to, i, j)
            a, b, c, d, i, j = args
            return bool(ctx.collide_pairs([i], [j], [a], [b], [c], [d])[0])
        if len(args) == 4:    # f(q_i, q_j, i, j)
            a, b, i, j = args
            return bool(ctx.collide_static_pairs([i], [j], [a], [b])[0])
        if len(args) == 2 and _is_config(args[0]):   # f(Q_from, Q_to)
            v, _ = ctx.collide_configs(_states(args[0], ctx.d), _states(args[1], ctx.d))
            return bool(v[0])
        if len(args) == 3 and _is_config(args[0]):   # f(Q, q_to, i)
            return bool(ctx.collide_config_moves(args[2], args[0], [args[1]])[0])
        raise TypeError("no collide method matches the arguments")

    f.ctx = ctx
    f.pairs = ctx.collide_pairs
    f.static_pairs = ctx.collide_static_pairs
    f.configs = ctx.collide_configs
    f.config_moves = ctx.collide_config_moves
    return f


def gen_check_goal(config_goal: Sequence, goal_rad: float = 0.0,
                   goal_rads: Optional[Sequence[float]] = None, dist=None,
                   ctx: Optional["Context"] = None):
    """utils.jl:117-138.  O(N) host logic (stays on the host in the reference design too:
    SURVEY 8(b)).  The reference tests `dist(q, goal) <= goal_rads[i]` with the MODEL's dist
    (utils.jl:126-134: arm22 wraps angles, line2d / snake2d / capsule3d weigh their components):
    pass `dist`, or `ctx` to use the model metric of the library (mrmp_state_dist).  Without
    either the test is the exact equality `q == goal` that goal_rad = 0 means for every model;
    a positive goal radius on a non-point model then needs `dist` or `ctx` and raises otherwise."""
    N = len(config_goal)
    rads = list(goal_rads) if goal_rads is not None else [goal_rad] * N
    point = isinstance(config_goal[0], (StatePoint2D, StatePoint3D)) if N else True
    if dist is None and ctx is None and not point and any(r > 0 for r in rads):
        raise ValueError("gen_check_goal: goal_rad > 0 on a non-point model needs the model's dist "
                         "(pass dist=... or ctx=connect.ctx)")

    def _d(a, b):
        if dist is not None:
            return dist(a, b)
        if ctx is not None:
            return float(ctx.state_dist([a], [b])[0])
        return math.sqrt(sum((x - y) * (x - y) for x, y in zip(a, b)))

    def _q(v):
        return v.q if isinstance(v, Node) else v

    def f(*args) -> bool:
        if len(args) == 1:   # f(C) / f(Q)
            return all(_d(_q(args[0][i]), config_goal[i]) <= rads[i] for i in range(N))
        v, i = args          # f(v, i)
        return _d(_q(v), config_goal[i]) <= rads[i]

    return f


def PRMs(config_init, config_goal, connect, num_vertices: int = 100, *, rad=None, rads=None,
         seed: int = 0) -> List[List[Node]]:
    """prm.jl:251-276 on the GPU: sampling (counter-based stream `seed`) + neighbour pass for
    every agent in one batched call.  `connect` must come from gen_connect()."""
    ctx: Context = connect.ctx
    N = len(config_init)
    rm = Roadmaps(ctx, 0, N, max(2, num_vertices))
    rm.sample(num_vertices, seed, config_init, config_goal, tries=False)
    rm.build(rads if rads is not None else rad)
    out = rm.to_nodes()
    rm.close()
    return out
