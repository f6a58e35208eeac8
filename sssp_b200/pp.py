"""Prioritized planning on the batched GPU conflict table.

Mirror of `MRMP.Solvers.PP` (src/solvers/pp.jl:46-137 on fresh roadmaps, :140-260 on given
roadmaps) and of the space-time A* it calls (`find_timed_path`, src/solvers/utils.jl:60-128),
with the reference's signatures and return value `(solution, roadmaps)`.  What runs where:

  GPU, once per planned agent i (instead of one `collide` call per generated search node):
      * the conflict table  bit[(v -> u), (t, j)] = collide(v.q, u.q, p_j[t-1], p_j[t], i, j)
        for every move of roadmap i (its edges plus the "stay" moves v -> v) against every
        already planned agent j at every time step t <= max_makespan -- two cross-product
        launches (`mrmp_collide_cross`: all moves with the OR over j taken on the device, and
        the stay moves per (t, j)); `invalid()` (pp.jl:171-193) and the extra goal constraints
        (pp.jl:195-213) become table look-ups;
      * the move costs dist(v.q, u.q) of the same moves (`mrmp_state_dist`);
      * beforehand, for all agents: the roadmaps (`PRMs`) and the distance tables with
        gen_g_func's edge cost (`mrmp_distance_tables_g`, pp.jl:161).
  host (as in the reference: branchy, sequential): the priority loop and the A* search with
      DataStructures.PriorityQueue's tie-breaking.

Ids and time steps are kept as in the reference's search (t is 1-based: the start is t = 1)
except vertex ids, which are 0-based.  `randperm(N)` of Julia's global RNG cannot be reproduced:
pass `order` explicitly, or `order_randomize=True` draws it from numpy's generator (`seed`).
"""
from __future__ import annotations

import math
import time
from typing import Callable, List, Optional, Sequence

import numpy as np

from .closures import Context, Node, PRMs, _states
from .solvers import PriorityQueue, get_distance_tables


def _egal_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """`q_from == q_to` of two immutable states: every field bit-identical."""
    return (a.view(np.uint64) == b.view(np.uint64)).all(axis=1)


class _AgentTables:
    """Everything `find_timed_path` asks about agent i, precomputed in two GPU calls."""

    def __init__(self, ctx: Context, roadmaps, i: int, paths, max_makespan: int,
                 stay_penalty: float, greedy: bool):
        rm = roadmaps[i]
        nv = len(rm)
        q = _states(rm, ctx.d)
        # moves of agent i in expansion order vcat(v.neighbors, v.id) (solvers/utils.jl:107)
        self.succ = [list(v.neighbors) + [v.id] for v in rm]
        self.row0 = np.zeros(nv + 1, np.int64)
        np.cumsum([len(s) for s in self.succ], out=self.row0[1:])
        src = np.repeat(np.arange(nv), np.diff(self.row0))
        dst = np.fromiter((u for s in self.succ for u in s), np.int64, count=int(self.row0[-1]))
        rf, rt = q[src], q[dst]
        # g_func(q_from, q_to, i) = dist + (q_from == q_to ? stay_penalty : 0)
        if greedy:
            self.cost = [0.0] * dst.size
        else:
            d = ctx.state_dist(rf, rt)
            self.cost = (d + np.where(_egal_rows(rf, rt), stay_penalty, 0.0)).tolist()
        # conflict table against the planned agents
        planned = [j for j in range(len(roadmaps)) if j != i and paths[j]]
        T = max_makespan - 1                       # time steps t = 2 .. max_makespan
        self.blocked = None                         # [move][t - 2]
        self.goal_blocked = None                    # [v][t - 1] for S.t = 1 .. max_makespan
        if not planned or T <= 0:
            return
        K = len(planned)
        ca = np.repeat(np.asarray(planned, np.int32)[None, :], T, axis=0).ravel()
        cf = np.empty((T, K, ctx.d))
        ct = np.empty((T, K, ctx.d))
        live = np.zeros((T, K), bool)               # t <= length(paths[j]): tested by check_goal
        for k, j in enumerate(planned):
            qj = _states(roadmaps[j], ctx.d)
            p = np.asarray(paths[j], np.int64)
            l = p.size
            t = np.arange(2, max_makespan + 1)
            cf[:, k] = qj[p[np.minimum(t - 1, l) - 1]]   # paths[j][min(t - 1, l)]
            ct[:, k] = qj[p[np.minimum(t, l) - 1]]       # paths[j][min(t, l)]
            live[:, k] = t <= l
        cf, ct = cf.reshape(-1, ctx.d), ct.reshape(-1, ctx.d)
        # invalid() (pp.jl:171-193) tests every planned agent at time t: OR over the K columns
        # of a time step, reduced on the device (group_size = K)
        bits = ctx.collide_cross(np.full(dst.size, i, np.int32), rf, rt, ca, cf, ct, group_size=K)
        self.blocked = np.unpackbits(bits.view(np.uint8), axis=1,
                                     bitorder="little")[:, :T].astype(bool)
        # check_goal_i (pp.jl:195-213): the stay move of v against paths[j][t-1 -> t] for
        # t = S.t + 1 .. length(paths[j]) only -> per-agent bits of the nv stay moves
        sbits = ctx.collide_cross(np.full(nv, i, np.int32), q, q, ca, cf, ct)
        hit = np.unpackbits(sbits.view(np.uint8), axis=1, bitorder="little")[:, :T * K]
        g = (hit.reshape(nv, T, K).astype(bool) & live[None, :, :]).any(axis=2)   # [v][t - 2]
        later = np.zeros((nv, max_makespan + 1), bool)             # index S.t - 1
        # later[v][s - 1] = any g[v][t - 2] for t >= s + 1
        acc = np.zeros(nv, bool)
        for s in range(max_makespan - 1, 0, -1):                   # S.t = s, first t = s + 1
            acc = acc | g[:, s - 1]
            later[:, s - 1] = acc
        self.goal_blocked = later


def find_timed_path(roadmap: Sequence[Node], tab: _AgentTables, check_goal_v: Callable,
                    h_table: Sequence[float], costs_table: List[float], max_makespan: int,
                    deadline: Optional[float] = None) -> Optional[List[int]]:
    """Space-time A* (solvers/utils.jl:60-128) with PP's invalid / check_goal_i / g_func_i
    (pp.jl:171-228) answered from the precomputed tables.  Returns vertex ids or None."""
    CLOSE = {}
    OPEN = PriorityQueue()
    num_generated = 1
    h0 = h_table[0]
    if h0 == math.inf:                               # case: not connected
        return None
    # search node = (v, t, parent_key, g, unique_id); its id "v-t" is the tuple (v, t)
    OPEN.enqueue((0, 1, None, 0.0, num_generated), 0.0 + h0)
    n_costs = len(costs_table)
    while len(OPEN):
        if deadline is not None and time.monotonic() > deadline:
            return None
        S = OPEN.dequeue()
        v, t, parent, g, _ = S
        key = (v, t)
        if key in CLOSE:
            continue
        CLOSE[key] = S
        # check_goal_i
        if check_goal_v(v) and not (tab.goal_blocked is not None and t <= max_makespan
                                    and tab.goal_blocked[v][t - 1]):
            path = []
            while S[2] is not None:
                path.insert(0, S[0])
                S = CLOSE[S[2]]
            path.insert(0, 0)
            return path
        # expand
        cost_to_come = max(g, costs_table[t - 1] if t <= n_costs else 0)   # pp.jl:220-221
        floor_next = costs_table[t] if t + 1 <= n_costs else 0
        r0 = int(tab.row0[v])
        for m, u in enumerate(tab.succ[v]):
            num_generated += 1
            g_new = max(cost_to_come + tab.cost[r0 + m], floor_next)       # pp.jl:222-225
            if (u, t + 1) in CLOSE:
                continue
            if t + 1 > max_makespan:                                       # pp.jl:176-178
                continue
            if tab.blocked is not None and tab.blocked[r0 + m][t - 1]:     # column t + 1
                continue
            OPEN.enqueue((u, t + 1, key, g_new, num_generated), g_new + h_table[u])
    return None


def PP_on_roadmaps(roadmaps, collide: Callable, check_goal: Callable,
                   g_func: Optional[Callable] = None, *, stay_penalty: float = 0.1,
                   greedy: bool = False, max_makespan: int = 20, order_randomize: bool = True,
                   order: Optional[Sequence[int]] = None, seed: int = 0,
                   TIME_LIMIT: Optional[float] = None, stats: Optional[dict] = None):
    """PP(roadmaps, collide, check_goal, g_func; ...) -> (solution | None, roadmaps), pp.jl:140-260.
    The cost function is gen_g_func(greedy, stay_penalty) (the only family the reference passes);
    a `g_func` made by solvers.gen_g_func overrides `greedy`."""
    ctx: Context = collide.ctx
    deadline = None if TIME_LIMIT is None else time.monotonic() + TIME_LIMIT
    if g_func is not None and getattr(g_func, "greedy", False):
        greedy = True
    N = len(roadmaps)
    tables = get_distance_tables(ctx, roadmaps, 1, 0.0 if greedy else stay_penalty)
    if greedy:
        tables = [[0.0 if math.isfinite(x) else x for x in t] for t in tables]
    paths: List[List[int]] = [[] for _ in range(N)]
    costs_table = [0.0]
    if order is None:
        order = np.random.default_rng(seed).permutation(N).tolist() if order_randomize else range(N)
    for i in order:
        if deadline is not None and time.monotonic() > deadline:
            return None, roadmaps
        t0 = time.perf_counter()
        tab = _AgentTables(ctx, roadmaps, i, paths, max_makespan, stay_penalty, greedy)
        t1 = time.perf_counter()
        rm = roadmaps[i]
        path = find_timed_path(rm, tab, lambda v, i=i, rm=rm: check_goal(rm[v], i), tables[i],
                               costs_table, max_makespan, deadline)
        if stats is not None:
            stats.setdefault("table_s", 0.0)
            stats.setdefault("search_s", 0.0)
            stats["table_s"] += t1 - t0
            stats["search_s"] += time.perf_counter() - t1
        if path is None:
            return None, roadmaps
        paths[i] = path
        # update costs table (pp.jl:245-256)
        r0 = tab.row0
        for t in range(2, len(path) + 1):
            v, u = path[t - 2], path[t - 1]
            c = tab.cost[int(r0[v]) + tab.succ[v].index(u)]
            c_new = max(costs_table[t - 1] if t <= len(costs_table) else 0, costs_table[t - 2] + c)
            if t <= len(costs_table):
                costs_table[t - 1] = c_new
            else:
                costs_table.append(c_new)
    # convert_paths_to_configurations (solvers/utils.jl:138-144)
    max_len = max(len(p) for p in paths)
    solution = [[roadmaps[i][paths[i][min(t, len(paths[i])) - 1]] for i in range(N)]
                for t in range(1, max_len + 1)]
    return solution, roadmaps


def _grow(roadmaps, fresh):
    """PRMs! on existing roadmaps (prm.jl:178-213): new vertices are appended and the all-pairs
    neighbour pass runs over the WHOLE roadmap again, pushing onto the old neighbour lists."""
    out = []
    for old, new in zip(roadmaps, fresh):
        rm = [Node(v.q, v.id, list(v.neighbors)) for v in new]
        for v in old:
            rm[v.id].neighbors = list(v.neighbors) + rm[v.id].neighbors
        out.append(rm)
    return out


def PP(config_init: Sequence, config_goal: Sequence, connect: Callable, collide: Callable,
       check_goal: Callable, *, g_func: Optional[Callable] = None, stay_penalty: float = 0.1,
       num_vertices: int = 100, rad=None, rads=None,
       roadmaps_growing_rate: Optional[float] = None, max_makespan: int = 20,
       order_randomize: bool = True, order: Optional[Sequence[int]] = None, seed: int = 0,
       TIME_LIMIT: Optional[float] = None, stats: Optional[dict] = None):
    """PP(config_init, config_goal, connect, collide, check_goal; ...) -> (solution | None,
    roadmaps), pp.jl:46-137: PRMs, then PP on them; on failure the roadmaps grow by
    `roadmaps_growing_rate` (the sampler streams continue, so the old vertices are kept)."""
    t_s = time.monotonic()

    def left():
        return None if TIME_LIMIT is None else TIME_LIMIT - (time.monotonic() - t_s)

    def timeover():
        return TIME_LIMIT is not None and time.monotonic() - t_s > TIME_LIMIT

    rng = np.random.default_rng(seed)

    def kw():
        # a fresh randperm(N) for every attempt (pp.jl:165), unless the caller fixed the order
        o = order
        if o is None:
            o = rng.permutation(len(config_init)).tolist() if order_randomize else None
        return dict(stay_penalty=stay_penalty, max_makespan=max_makespan, order_randomize=False,
                    order=o, stats=stats)
    roadmaps = PRMs(config_init, config_goal, connect, num_vertices, rad=rad, rads=rads, seed=seed)
    if timeover():
        return None, roadmaps
    solution, roadmaps = PP_on_roadmaps(roadmaps, collide, check_goal, g_func, TIME_LIMIT=left(),
                                        **kw())
    while solution is None and roadmaps_growing_rate is not None and not timeover():
        num_vertices = int(math.floor(num_vertices * roadmaps_growing_rate))
        fresh = PRMs(config_init, config_goal, connect, num_vertices, rad=rad, rads=rads,
                     seed=seed)
        roadmaps = _grow(roadmaps, fresh)
        if timeover():
            break
        if stats is not None:
            stats["attempts"] = stats.get("attempts", 1) + 1
        solution, roadmaps = PP_on_roadmaps(roadmaps, collide, check_goal, g_func,
                                            TIME_LIMIT=left(), **kw())
    return solution, roadmaps
