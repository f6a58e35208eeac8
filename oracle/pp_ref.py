"""CPU ORACLE (test infrastructure, NOT product code): line-by-line restatement of the
reference's prioritized planning over the oracle closures -- one `collide` call per generated
search node, exactly as the Julia code issues them.

Follows /root/reference/src/solvers/pp.jl:140-260 (PP on given roadmaps: invalid :171-193,
check_goal_i :195-213, g_func_i :217-228, costs table :245-256), src/solvers/utils.jl:29-39
(gen_g_func), :60-128 (find_timed_path), :138-144 (convert_paths_to_configurations) and
:146-179 (get_distance_table with g_func).  Vertex ids and agent indices are 0-based, time
steps 1-based as in the reference.

Outside /root/reference (parity unpinned): DataStructures.PriorityQueue's
tie-breaking (class _PQ of sssp_ref.py) and `randperm` (the caller passes the order).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from .sssp_ref import Node, _PQ


def gen_g_func(orc, greedy=False, stay_penalty=0.0):
    """solvers/utils.jl:34-36: f(q_from, q_to, i = 1)."""
    def f(q_from, q_to):
        if greedy:
            return 0.0
        same = np.array_equal(np.asarray(q_from).view(np.uint64), np.asarray(q_to).view(np.uint64))
        return orc.state_dist(q_from, q_to) + (stay_penalty if same else 0)
    return f


def get_distance_table(roadmap: List[Node], g_func, goal: int = 1) -> List[float]:
    """solvers/utils.jl:146-179."""
    table = [math.inf] * len(roadmap)
    OPEN = _PQ()
    table[goal] = 0.0
    OPEN.enqueue(goal, 0.0)
    while len(OPEN):
        v_id = OPEN.dequeue()
        v = roadmap[v_id]
        d = table[v_id]
        for u_id in v.neighbors:
            g = g_func(v.q, roadmap[u_id].q) + d
            if g < table[u_id]:
                if u_id in OPEN:
                    OPEN.delete(u_id)
                table[u_id] = g
                OPEN.enqueue(u_id, g)
    return table


def find_timed_path(roadmap, invalid, check_goal, h_func, g_func, max_pops=2_000_000):
    """solvers/utils.jl:60-128.  A search node is a dict; its id is "v-t"."""
    CLOSE = {}
    OPEN = _PQ()
    num_generated_nodes = 1
    v_init = roadmap[0]
    S_init = dict(v=v_init, t=1, id="%d-%d" % (v_init.id, 1), parent_id=None, g=0.0,
                  h=h_func(v_init), unique_id=num_generated_nodes)
    S_init["f"] = S_init["g"] + S_init["h"]
    if S_init["h"] == math.inf:
        return None
    nodes = {1: S_init}
    OPEN.enqueue(1, S_init["f"])
    pops = 0
    while len(OPEN) and pops < max_pops:
        pops += 1
        S = nodes.pop(OPEN.dequeue())
        if S["id"] in CLOSE:
            continue
        CLOSE[S["id"]] = S
        if check_goal(S):
            path = []
            while S["parent_id"] is not None:
                path.insert(0, S["v"])
                S = CLOSE[S["parent_id"]]
            path.insert(0, v_init)
            return path
        for k in list(S["v"].neighbors) + [S["v"].id]:
            u = roadmap[k]
            num_generated_nodes += 1
            S_new = dict(v=u, t=S["t"] + 1, id="%d-%d" % (u.id, S["t"] + 1), parent_id=S["id"],
                         g=g_func(S, u.q), h=h_func(u), unique_id=num_generated_nodes)
            S_new["f"] = S_new["g"] + S_new["h"]
            if S_new["id"] in CLOSE or invalid(S, S_new):
                continue
            nodes[num_generated_nodes] = S_new
            OPEN.enqueue(num_generated_nodes, S_new["f"])
    return None


def PP(orc, roadmaps: List[List[Node]], config_goal, *, goal_rads=None, stay_penalty=0.1,
       greedy=False, max_makespan=20, order: Optional[Sequence[int]] = None, counters=None):
    """pp.jl:140-260.  Returns (solution | None, paths); solution[t][i] = vertex id."""
    N = len(roadmaps)
    goal_rads = [0.0] * N if goal_rads is None else list(goal_rads)
    config_goal = [np.asarray(q, float) for q in config_goal]
    g_func = gen_g_func(orc, greedy, stay_penalty)
    distance_tables = [get_distance_table(rm, g_func) for rm in roadmaps]
    paths: List[List[Node]] = [[] for _ in range(N)]
    costs_table = [0.0]
    n_collide = 0
    for i in (range(N) if order is None else order):

        def invalid(S_from, S_to, i=i):
            nonlocal n_collide
            t = S_to["t"]
            if t > max_makespan:
                return True
            for j in range(N):
                if j == i or not paths[j]:
                    continue
                l = len(paths[j])
                q_j_from = paths[j][min(t - 1, l) - 1].q
                q_j_to = paths[j][min(t, l) - 1].q
                n_collide += 1
                if orc.collide_pair(S_from["v"].q, S_to["v"].q, q_j_from, q_j_to, i, j):
                    return True
            return False

        def check_goal_i(S, i=i):
            nonlocal n_collide
            if not orc.state_dist(S["v"].q, config_goal[i]) <= goal_rads[i]:   # utils.jl:132
                return False
            for j in range(N):
                if j == i or not paths[j]:
                    continue
                for t in range(S["t"] + 1, len(paths[j]) + 1):
                    n_collide += 1
                    if orc.collide_pair(S["v"].q, S["v"].q, paths[j][t - 2].q, paths[j][t - 1].q,
                                        i, j):
                        return False
            return True

        def h_func_i(v, i=i):
            return distance_tables[i][v.id]

        def g_func_i(S, q):
            t = S["t"]
            cost_to_come_S = max(S["g"], costs_table[t - 1] if t <= len(costs_table) else 0)
            return max(cost_to_come_S + g_func(S["v"].q, q),
                       costs_table[t] if t + 1 <= len(costs_table) else 0)

        path = find_timed_path(roadmaps[i], invalid, check_goal_i, h_func_i, g_func_i)
        if path is None:
            if counters is not None:
                counters["collide_calls"] = n_collide
            return None, paths
        paths[i] = path
        for t in range(2, len(path) + 1):
            c_new = max(costs_table[t - 1] if t <= len(costs_table) else 0,
                        costs_table[t - 2] + g_func(path[t - 2].q, path[t - 1].q))
            if t <= len(costs_table):
                costs_table[t - 1] = c_new
            else:
                costs_table.append(c_new)
    if counters is not None:
        counters["collide_calls"] = n_collide
    max_len = max(len(p) for p in paths)
    solution = [[paths[i][min(t, len(paths[i])) - 1].id for i in range(N)]
                for t in range(1, max_len + 1)]
    return solution, paths
