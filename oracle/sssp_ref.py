"""CPU ORACLE (test infrastructure, NOT product code): line-by-line restatement of the
reference's SSSP solver over the oracle closures.

Follows /root/reference/src/solvers/sssp.jl (SSSP :49-232, expand! :235-267, steering
:269-291, gen_RRT_connect_roadmaps :294-401, extend! :404-440, get_Q_id :443-445, backtrack
:448-461) and src/solvers/utils.jl:146-179 (get_distance_table).  Vertex ids and agent
indices are 0-based here.

Arithmetic that lives outside /root/reference (parity unpinned):
  * DataStructures.PriorityQueue (un-pinned dependency, Project.toml): binary min-heap with
    percolate_up!/percolate_down! comparing priorities with `<` only -- its tie-breaking decides
    which of several equal-f search nodes is popped first.  Restated below (class _PQ).
  * the sampler: Julia's global RNG cannot be reproduced, so sampler() is the counter-based
    stream orc_sample_state(seed, agent, t) with one counter per agent, consumed in call order.
  * sum(...)/N of the heuristic: accumulated left to right.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np


class _PQ:
    """DataStructures.PriorityQueue{K,Float64} (binary heap, Base.Order.Forward)."""

    def __init__(self):
        self.xs: list = []          # (key, priority)
        self.index: dict = {}

    def __len__(self):
        return len(self.xs)

    def __contains__(self, key):
        return key in self.index

    def _up(self, i):
        xs = self.xs
        x = xs[i]
        while i > 0:
            j = (i - 1) // 2
            xj = xs[j]
            if x[1] < xj[1]:
                self.index[xj[0]] = i
                xs[i] = xj
            else:
                break
            i = j
        self.index[x[0]] = i
        xs[i] = x

    def _down(self, i):
        xs = self.xs
        n = len(xs)
        x = xs[i]
        while True:
            l = 2 * i + 1
            if l >= n:
                break
            r = l + 1
            j = l if (r >= n or xs[l][1] < xs[r][1]) else r
            xj = xs[j]
            if xj[1] < x[1]:
                self.index[xj[0]] = i
                xs[i] = xj
            else:
                break
            i = j
        self.index[x[0]] = i
        xs[i] = x

    def enqueue(self, key, pri):
        self.xs.append((key, pri))
        self.index[key] = len(self.xs) - 1
        self._up(len(self.xs) - 1)

    def dequeue(self):
        x = self.xs[0]
        y = self.xs.pop()
        if self.xs:
            self.xs[0] = y
            self.index[y[0]] = 0
            self._down(0)
        del self.index[x[0]]
        return x[0]

    def delete(self, key):
        # DataStructures.delete!: move to the root with priority -inf semantics, then dequeue
        i = self.index[key]
        xs = self.xs
        x = xs[i]
        while i > 0:
            j = (i - 1) // 2
            self.index[xs[j][0]] = i
            xs[i] = xs[j]
            i = j
        xs[0] = x
        self.index[key] = 0
        self.dequeue()


class Node:
    __slots__ = ("q", "id", "neighbors")

    def __init__(self, q, id, neighbors=None):
        self.q, self.id, self.neighbors = np.asarray(q, dtype=np.float64), id, neighbors or []


class Sampler:
    """sampler() of agent i = next try of the counter-based stream (seed, i)."""

    def __init__(self, orc, seed, N):
        self.orc, self.seed, self.t = orc, seed, [0] * N

    def __call__(self, i):
        q = self.orc.sample_state(self.seed, i, self.t[i])
        self.t[i] += 1
        return q


def get_distance_table(orc, roadmap: List[Node], goal: int = 1) -> List[float]:
    """solvers/utils.jl:146-179 (g_func = dist)."""
    table = [math.inf] * len(roadmap)
    OPEN = _PQ()
    table[goal] = 0.0
    OPEN.enqueue(goal, 0.0)
    while len(OPEN):
        v_id = OPEN.dequeue()
        v = roadmap[v_id]
        d = table[v_id]
        for u_id in v.neighbors:
            g = orc.state_dist(v.q, roadmap[u_id].q) + d
            if g < table[u_id]:
                if u_id in OPEN:
                    OPEN.delete(u_id)
                table[u_id] = g
                OPEN.enqueue(u_id, g)
    return table


def steering(orc, conn, q_rand, q_near, depth):
    """sssp.jl:269-291"""
    q_h = q_rand
    if not conn(q_near, q_h):
        q_l = q_near
        for _ in range(depth):
            q = orc.mid_status(q_l, q_h)
            if conn(q_near, q):
                q_l = q
            else:
                q_h = q
        q_h = q_l
    return q_h


def expand(orc, conn, sample, v_from: Node, roadmap: List[Node], min_dist_thread, n_expansion,
           depth) -> bool:
    """sssp.jl:235-267"""
    updated = False
    for _ in range(n_expansion):
        q_new = steering(orc, conn, sample(), v_from.q, depth)
        if min(orc.state_dist(v.q, q_new) for v in roadmap) > min_dist_thread:
            u = Node(q_new, len(roadmap), [v.id for v in roadmap if conn(q_new, v.q)])
            roadmap.append(u)
            for v in roadmap:
                if conn(v.q, q_new):
                    v.neighbors.append(u.id)
            updated = True
    return updated


def _extend(orc, conn2, conn1, q_rand, V, from_start, depth):
    """sssp.jl:404-440"""
    if from_start:
        dists = [orc.state_dist(q, q_rand) for q in V]
    else:
        dists = [orc.state_dist(q_rand, q) for q in V]
    q_near = V[int(np.argmin(dists))]      # first minimum, like Julia's argmin
    flg = "reached"
    q_l, q_h = q_near, q_rand
    bad = (not conn2(q_near, q_h)) if from_start else (not (conn1(q_h) and conn2(q_h, q_near)))
    if bad:
        flg = "trapped"
        for _ in range(depth):
            q = orc.mid_status(q_l, q_h) if from_start else orc.mid_status(q_h, q_l)
            ok = conn2(q_near, q) if from_start else (conn1(q) and conn2(q, q_near))
            if ok:
                flg = "advanced"
                q_l = q
            else:
                q_h = q
        q_h = q_l
    if flg != "trapped":
        V.append(q_h)
    return flg, q_h


def gen_RRT_connect_roadmap(orc, i, q_init, q_goal, sample, depth, epsilon, max_iter=100000):
    """sssp.jl:325-401 (the wall-clock limit is replaced by an iteration cap)."""
    def conn2(qf, qt):
        if not (epsilon is None or orc.state_dist(qf, qt) <= epsilon):
            return False
        return orc.connect_edge(qf, qt, i)

    def conn1(q):
        return orc.connect_point(q, i)

    q_init, q_goal = np.asarray(q_init, float), np.asarray(q_goal, float)
    if conn2(q_init, q_goal):
        v_init, v_goal = Node(q_init, 0, [1]), Node(q_goal, 1, [])
        if conn2(q_goal, q_init):
            v_goal.neighbors.append(0)
        return [v_init, v_goal]
    V1, V2 = [q_init], [q_goal]
    it = 0
    while it < max_iter:
        it += 1
        from_start = it % 2 == 1
        flg, q_new = _extend(orc, conn2, conn1, sample(), V1, from_start, depth)
        if flg != "trapped" and _extend(orc, conn2, conn1, q_new, V2, not from_start, depth)[0] == "reached":
            if from_start:
                V1, V2 = V2, V1
            roadmap = [Node(q_init, 0), Node(q_goal, 1)]
            for q in V1[1:]:
                roadmap.append(Node(q, len(roadmap)))
            for q in V2[1:]:
                roadmap.append(Node(q, len(roadmap)))
            for k, vf in enumerate(roadmap):
                for l, vt in enumerate(roadmap):
                    if k != l and conn2(vf.q, vt.q):
                        vf.neighbors.append(vt.id)
            return roadmap
        V1, V2 = V2, V1
    return None


def get_Q_id(Q: List[Node], nxt: int) -> str:
    return "-".join(str(v.id) for v in Q) + "_%d" % nxt


def SSSP(orc, config_init, config_goal, *, goal_rads=None, steering_depth=2,
         num_vertex_expansion=10, init_min_dist_thread=0.1, decreasing_rate_min_dist_thread=0.99,
         epsilon=None, no_roadmap_at_beginning=False, no_fast_collision_check=False, seed=0,
         max_rounds=50, max_loops=200000):
    """sssp.jl:49-232 with g_func = greedy (the default).  Returns (solution, roadmaps, stats);
    solution = list of configurations, each a list of vertex ids (one per agent)."""
    N = len(config_init)
    config_init = [np.asarray(q, float) for q in config_init]
    config_goal = [np.asarray(q, float) for q in config_goal]
    goal_rads = [0.0] * N if goal_rads is None else list(goal_rads)
    smp = Sampler(orc, seed, N)

    def conn_i(i):
        def f(qf, qt):
            return (epsilon is None or orc.state_dist(qf, qt) <= epsilon) and orc.connect_edge(qf, qt, i)
        return f

    roadmaps = []
    for i in range(N):
        if no_roadmap_at_beginning:
            v_init, v_goal = Node(config_init[i], 0), Node(config_goal[i], 1)
            if conn_i(i)(v_init.q, v_goal.q):
                v_init.neighbors.append(1)
            if conn_i(i)(v_goal.q, v_init.q):
                v_goal.neighbors.append(0)
            roadmaps.append([v_init, v_goal])
        else:
            rmp = gen_RRT_connect_roadmap(orc, i, config_init[i], config_goal[i], lambda i=i: smp(i),
                                          steering_depth, epsilon)
            if rmp is None:
                return None, roadmaps, {"expansions": 0}
            roadmaps.append(rmp)
    tables = [get_distance_table(orc, rmp) for rmp in roadmaps]

    def h_func(Q):
        s = 0.0
        for i in range(N):
            s = s + tables[i][Q[i].id]
        return s / N

    def check_goal(Q):
        return all(orc.state_dist(Q[i].q, config_goal[i]) <= goal_rads[i] for i in range(N))

    Q_init = [roadmaps[i][0] for i in range(N)]
    S_init = dict(Q=Q_init, next=0, id=get_Q_id(Q_init, -1), parent=None, g=0.0, h=h_func(Q_init),
                  depth=1)
    S_init["f"] = S_init["g"] + S_init["h"]
    n_exp = 0
    for k in range(1, max_rounds + 1):
        OPEN = _PQ()
        VISITED = {}
        min_dist_thread = init_min_dist_thread * (decreasing_rate_min_dist_thread ** (k - 1))
        OPEN.enqueue(S_init["id"], S_init["f"])
        VISITED[S_init["id"]] = S_init
        loops = 0
        while len(OPEN) and loops < max_loops:
            loops += 1
            S = VISITED[OPEN.dequeue()]
            if check_goal(S["Q"]):
                sol = []
                T = S
                while T["parent"] is not None:
                    sol.insert(0, [v.id for v in T["Q"]])
                    T = VISITED[T["parent"]]
                sol.insert(0, [v.id for v in T["Q"]])
                return sol, roadmaps, {"expansions": n_exp, "rounds": k}
            i = S["next"]
            j = (i + 1) % N
            v = S["Q"][i]
            n_exp += 1
            if expand(orc, conn_i(i), lambda: smp(i), v, roadmaps[i], min_dist_thread,
                      num_vertex_expansion, steering_depth):
                tables[i] = get_distance_table(orc, roadmaps[i])
            Qs = np.stack([u.q for u in S["Q"]])
            for p_id in list(v.neighbors) + [v.id]:
                p = roadmaps[i][p_id]
                Q = list(S["Q"])
                Q[i] = p
                Q_id = get_Q_id(Q, j)
                if Q_id in VISITED:
                    continue
                if not no_fast_collision_check:
                    if orc.collide_config_move(Qs, p.q, i):
                        continue
                else:
                    Qt = Qs.copy()
                    Qt[i] = p.q
                    if orc.collide_configs(Qs, Qt)[0]:
                        continue
                S_new = dict(Q=Q, next=j, id=Q_id, parent=S["id"], h=h_func(Q), g=S["g"] + 0.0,
                             depth=S["depth"] + 1)
                S_new["f"] = S_new["g"] + S_new["h"]
                OPEN.enqueue(S_new["id"], S_new["f"])
                VISITED[S_new["id"]] = S_new
    return None, roadmaps, {"expansions": n_exp}
