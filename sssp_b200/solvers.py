"""Host side of the reference's SSSP solver on top of the batched GPU expansion.

Mirror of `MRMP.Solvers.SSSP` (src/solvers/sssp.jl:49-232) with the reference's signature and
return value `(solution, roadmaps)`.  What runs where:

  GPU (one `mrmp_sssp_expand` call per popped search node, roadmaps resident in HBM)
      expand! (:235-267): sampling, steering (:269-291), the space-filling test (:250), every
      conn(q_new, v.q) / conn(v.q, q_new) against the whole roadmap (:252-262), and the
      successor filter collide(S.Q, p.q, i) / collide(S.Q, Q) (:206-207);
      the all-pairs edge pass of the initial roadmaps (:388-395) = `mrmp_prm_build`;
      the distance tables (solvers/utils.jl:146-179) = `mrmp_distance_tables`.
  host (as in the reference: branchy, sequential)
      the best-first search: OPEN / VISITED, successor generation, backtracking (:146-232);
      RRT-Connect tree growth (:372-401) -- sequential by construction, each test one GPU call.

Ids are 0-based.  The reference draws samples from Julia's global RNG; here `sampler()` of
agent i is the next try of the library's counter-based stream (seed, i), so runs are
reproducible anywhere.  OPEN is DataStructures.PriorityQueue's binary heap restated (its
tie-breaking among equal f-values decides the search order).
"""
from __future__ import annotations

import math
import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import lib as L
from .closures import Context, Node, SsspRoadmaps, _STATE_OF, _states, collide_params_agree


class PriorityQueue:
    """Binary min-heap with DataStructures.PriorityQueue's percolate rules: a child moves up
    only on strictly smaller priority; dequeue moves the last element to the root."""

    def __init__(self):
        self._heap: List[Tuple[object, float]] = []
        self._pos: dict = {}

    def __len__(self):
        return len(self._heap)

    def __contains__(self, key):
        return key in self._pos

    def _place(self, i, item):
        self._heap[i] = item
        self._pos[item[0]] = i

    def _sift_up(self, i, force=False):
        item = self._heap[i]
        while i > 0:
            parent = (i - 1) >> 1
            if force or item[1] < self._heap[parent][1]:
                self._place(i, self._heap[parent])
                i = parent
            else:
                break
        self._place(i, item)

    def _sift_down(self, i):
        item, n = self._heap[i], len(self._heap)
        while 2 * i + 1 < n:
            left, right = 2 * i + 1, 2 * i + 2
            child = left if right >= n or self._heap[left][1] < self._heap[right][1] else right
            if self._heap[child][1] < item[1]:
                self._place(i, self._heap[child])
                i = child
            else:
                break
        self._place(i, item)

    def enqueue(self, key, priority):
        self._heap.append((key, priority))
        self._sift_up(len(self._heap) - 1)

    def dequeue(self):
        top, last = self._heap[0], self._heap.pop()
        if self._heap:
            self._heap[0] = last
            self._sift_down(0)
        del self._pos[top[0]]
        return top[0]

    def delete(self, key):
        self._sift_up(self._pos[key], force=True)
        self.dequeue()


def gen_g_func(greedy: bool = False, stay_penalty: float = 0.0, ctx: Optional[Context] = None):
    """solvers/utils.jl:29-39.  The non-greedy form needs the model's dist: pass the ctx."""
    if not greedy and ctx is None:
        raise ValueError("gen_g_func(greedy=False) needs ctx (the model's dist runs in the library)")

    def f(Q_from, Q_to) -> float:
        if greedy:
            return 0.0
        d = ctx.state_dist([v.q for v in Q_from], [v.q for v in Q_to])
        s = 0.0
        for x in d:
            s += float(x)
        return s / len(Q_from)
    f.greedy = greedy
    return f


def get_distance_table(ctx: Context, nodes: np.ndarray, nbrs: List[List[int]], goal: int = 1):
    """solvers/utils.jl:146-179: cost-to-go from the goal vertex along the neighbour lists.
    Runs on the GPU (mrmp_distance_tables: parallel edge relaxation to the fixed point, which
    equals the reference's Dijkstra bit for bit)."""
    return ctx.distance_tables([(nodes, nbrs)], goal)[0].tolist()


def get_distance_tables(ctx: Context, roadmaps, goal: int = 1, stay_penalty: float = 0.0):
    """solvers/utils.jl:226-236 for Vector{Vector{Node}} roadmaps: all agents in one launch."""
    graphs = [(np.array([tuple(v.q) for v in rm]), [list(v.neighbors) for v in rm]) for rm in roadmaps]
    return [t.tolist() for t in ctx.distance_tables(graphs, goal, stay_penalty)]


class _AgentSampler:
    def __init__(self, ctx: Context, seed: int, N: int):
        self.ctx, self.seed, self.t = ctx, seed, [0] * N

    def draw(self, i: int) -> np.ndarray:
        q = self.ctx.sample_states(i, self.seed, self.t[i], 1)[0]
        self.t[i] += 1
        return q


def _rrt_connect_roadmap(ctx: Context, i: int, q_init, q_goal, smp: _AgentSampler, depth: int,
                         epsilon: Optional[float], deadline: Optional[float]):
    """gen_RRT_connect_roadmap (sssp.jl:325-401) -> (nodes [nv][d], neighbour lists) or None."""
    ag = [i]

    def conn2(qf, qt) -> bool:
        if epsilon is not None and not float(ctx.state_dist([qf], [qt])[0]) <= epsilon:
            return False
        return bool(ctx.connect_edges(ag, [qf], [qt])[0])

    def conn1(q) -> bool:
        return bool(ctx.connect_points(ag, [q])[0])

    def extend(q_rand, V, from_start):          # extend!, sssp.jl:404-440: one GPU call
        flg, q_h, _ = ctx.rrt_extend(i, np.stack(V), q_rand, from_start, depth, epsilon)
        if flg != "trapped":
            V.append(q_h)
        return flg, q_h

    q_init, q_goal = np.asarray(q_init, float), np.asarray(q_goal, float)
    if conn2(q_init, q_goal):
        return np.stack([q_init, q_goal]), [[1], [0] if conn2(q_goal, q_init) else []]
    V1, V2 = [q_init], [q_goal]
    it = 0
    while deadline is None or time.monotonic() < deadline:
        it += 1
        from_start = it % 2 == 1
        flg, q_new = extend(smp.draw(i), V1, from_start)
        if flg != "trapped" and extend(q_new, V2, not from_start)[0] == "reached":
            if from_start:
                V1, V2 = V2, V1
            nodes = np.stack([q_init, q_goal] + V1[1:] + V2[1:])
            # all ordered pairs k != l with conn (sssp.jl:388-395): the PRM neighbour pass
            row_ptr, col = ctx.prm_build(nodes[None], epsilon, agent0=i)
            return nodes, [col[row_ptr[v]: row_ptr[v + 1]].tolist() for v in range(nodes.shape[0])]
        V1, V2 = V2, V1
    return None


def SSSP(config_init: Sequence, config_goal: Sequence, connect: Callable, collide: Callable,
         check_goal: Callable, *, g_func: Optional[Callable] = None, steering_depth: int = 2,
         num_vertex_expansion: int = 10, init_min_dist_thread: float = 0.1,
         decreasing_rate_min_dist_thread: float = 0.99, epsilon: Optional[float] = None,
         TIME_LIMIT: Optional[float] = 30, VERBOSE: int = 0, use_random_h_func: bool = False,
         no_roadmap_at_beginning: bool = False, no_fast_collision_check: bool = False,
         seed: int = 0, stats: Optional[dict] = None):
    """SSSP(config_init, config_goal, connect, collide, check_goal; kwargs...) ->
    (solution | None, roadmaps): sssp.jl:49-73.  `connect` / `collide` come from gen_connect /
    gen_collide (built from the same rads); `seed` selects the sampler streams."""
    ctx: Context = connect.ctx
    # The fused expansion runs the steering (connect) and the successor filter (collide) of one
    # popped node in ONE launch on connect's context: that is only the reference's computation
    # when the collide closure was built from the same parameters.
    if not collide_params_agree(collide.ctx, ctx):
        raise ValueError("SSSP: connect and collide must be built from the same ins_params, "
                         "step_dist and safety_dist (the node expansion evaluates both on one context)")
    t_s = time.monotonic()
    deadline = None if TIME_LIMIT is None else t_s + TIME_LIMIT

    def timeover() -> bool:
        return deadline is not None and time.monotonic() > deadline

    N, d = len(config_init), ctx.d
    State = _STATE_OF[ctx.model]
    g_func = g_func if g_func is not None else gen_g_func(greedy=True)
    init, goal = _states(config_init, d), _states(config_goal, d)
    smp = _AgentSampler(ctx, seed, N)
    rng_h = np.random.default_rng(seed) if use_random_h_func else None

    # ---- initial roadmaps (sssp.jl:91-112)
    tabs: List[np.ndarray] = []
    nbrs: List[List[List[int]]] = []
    for i in range(N):
        if no_roadmap_at_beginning:
            ok = ctx.connect_edges([i, i], [init[i], goal[i]], [goal[i], init[i]]).astype(bool)
            if epsilon is not None:
                ok &= ctx.state_dist([init[i], goal[i]], [goal[i], init[i]]) <= epsilon
            rm = (np.stack([init[i], goal[i]]), [[1] if ok[0] else [], [0] if ok[1] else []])
        else:
            rm = _rrt_connect_roadmap(ctx, i, init[i], goal[i], smp, steering_depth, epsilon,
                                      deadline)
            if rm is None:
                return None, [[] for _ in range(N)]
        tabs.append(rm[0])
        nbrs.append(rm[1])
    dev = SsspRoadmaps(ctx, nv_cap=max(64, 2 * max(t.shape[0] for t in tabs)))
    for i in range(N):
        dev.set_roadmap(i, tabs[i])
    tables = [t.tolist() for t in ctx.distance_tables(list(zip(tabs, nbrs)))]

    def h_func(ids) -> float:
        if use_random_h_func:
            return float(rng_h.random())
        s = 0.0
        for i in range(N):
            s = s + tables[i][ids[i]]
        return s / N

    def as_nodes(ids):
        return [Node(State(*tabs[i][ids[i]]), ids[i], nbrs[i][ids[i]]) for i in range(N)]

    def roadmaps_out():
        return [[Node(State(*tabs[i][v]), v, list(nbrs[i][v])) for v in range(tabs[i].shape[0])]
                for i in range(N)]

    Q0 = tuple(0 for _ in range(N))
    # search node: key (ids, next); record [parent_key, g, depth]
    S_init_key = (Q0, -1)
    n_expand = 0
    k = 0
    while not timeover():
        k += 1
        OPEN = PriorityQueue()
        VISITED = {S_init_key: (None, 0.0, 1, 0)}       # parent, g, depth, next
        min_dist_thread = init_min_dist_thread * (decreasing_rate_min_dist_thread ** (k - 1))
        OPEN.enqueue(S_init_key, 0.0 + h_func(Q0))
        while len(OPEN) and not timeover():
            key = OPEN.dequeue()
            ids = key[0]
            parent, g, depth, i = VISITED[key]
            if check_goal(as_nodes(ids)):
                sol, cur = [], key
                while cur is not None:
                    sol.insert(0, as_nodes(cur[0]))
                    cur = VISITED[cur][0]
                if stats is not None:
                    stats.update(expansions=n_expand, rounds=k, solution_ids=[
                        [v.id for v in Q] for Q in sol])
                dev.close()
                return sol, roadmaps_out()
            j = (i + 1) % N
            v = ids[i]
            # ---- GPU: expand! + successor filter in one call (sssp.jl:181-223)
            q_new, outs, ins, succ, hit = dev.expand(
                i, v, ids, nbrs[i][v], num_vertex_expansion, seed=seed, t0=smp.t[i],
                steering_depth=steering_depth, epsilon=epsilon, min_dist_thread=min_dist_thread,
                slow_collide=no_fast_collision_check)
            smp.t[i] += num_vertex_expansion
            n_expand += 1
            if len(q_new):
                base = tabs[i].shape[0]
                tabs[i] = np.concatenate([tabs[i], q_new])
                for r in range(len(q_new)):
                    nbrs[i].append(list(outs[r]))
                    for t in ins[r]:
                        nbrs[i][t].append(base + r)
                tables[i] = get_distance_table(ctx, tabs[i], nbrs[i])
            for p_id, blocked in zip(succ.tolist(), hit.tolist()):
                new_ids = ids[:i] + (p_id,) + ids[i + 1:]
                new_key = (new_ids, j)
                if new_key in VISITED or blocked:
                    continue
                g_new = g + g_func(as_nodes(ids), as_nodes(new_ids)) if not getattr(
                    g_func, "greedy", False) else g + 0.0
                VISITED[new_key] = (key, g_new, depth + 1, j)
                OPEN.enqueue(new_key, g_new + h_func(new_ids))
    if stats is not None:
        stats.update(expansions=n_expand, rounds=k, solution_ids=None)
    dev.close()
    return None, roadmaps_out()
