"""TEST INFRASTRUCTURE -- literal restatement of the reference's solution smoother
(/root/reference/src/smoother.jl) over the CPU oracle's closures: one collide / connect call per
check, in the reference's loop order.  Only tests and bench.py's CPU legs may import this.

Conventions: agents and time steps are 0-based here (the reference is 1-based); an action's `t`
is the 0-based index of the configuration in which the new vertex first appears, so ids read
"<from.id>_<to.id>_<t>" with t one less than in Julia.  Julia iterates `values(Dict)` in hash
order when it rebuilds the predecessor lists (smoother.jl:111-119); here they are listed by
ascending agent -- the SET of dependencies is what the reference defines.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Set, Tuple

import numpy as np


@dataclass
class Vertex:          # what an Action needs of the reference's Node
    q: np.ndarray
    id: int


@dataclass
class Action:          # smoother.jl:13-21
    id: str
    frm: Vertex
    to: Vertex
    t: int
    agent: int
    predecessors: List[Tuple[int, str]] = field(default_factory=list)
    successors: List[Tuple[int, str]] = field(default_factory=list)


def get_action_id(frm: Vertex, to: Vertex, t: int) -> str:   # smoother.jl:23-29
    return "%d_%d_%d" % (frm.id, to.id, t)


def _find(TPG, j, id_):
    for a in TPG[j]:
        if a.id == id_:
            return a
    raise KeyError((j, id_))


def get_temporal_plan_graph(solution, collide, connect, skip_connection: bool = True):
    """smoother.jl:32-132.  solution[t][i] = Vertex; collide(qf_i, qt_i, qf_j, qt_j, i, j),
    connect(q_from, q_to, i) are closures over the oracle."""
    N, T = len(solution[0]), len(solution)
    TPG: List[List[Action]] = [[] for _ in range(N)]
    for i in range(N):                                   # type-1 dependency (:45-75)
        v_last = solution[0][i]
        for t, Q in enumerate(solution):
            v = Q[i]
            if v.id != v_last.id:
                a = Action(get_action_id(v_last, v, t), v_last, v, t, i)
                if TPG[i]:
                    a.predecessors.append((i, TPG[i][-1].id))
                TPG[i].append(a)
                v_last = v
    for i in range(N):                                   # type-2 dependency (:78-106)
        for a_self in TPG[i]:
            for j in range(N):
                if j == i:
                    continue
                for a_other in [a for a in TPG[j] if a.t > a_self.t]:
                    if collide(a_self.frm.q, a_self.to.q, a_other.frm.q, a_other.to.q, i, j):
                        a_other.predecessors.append((a_self.agent, a_self.id))
                        break
    for i in range(N):                                   # remove redundant dependencies (:109-127)
        for a in TPG[i]:
            latest: Dict[int, Action] = {}
            for (j, id_) in a.predecessors:
                o = _find(TPG, j, id_)
                if j not in latest or latest[j].t < o.t:
                    latest[j] = o
            a.predecessors = [(latest[j].agent, latest[j].id) for j in sorted(latest)]
            for (j, id_) in a.predecessors:
                _find(TPG, j, id_).successors.append((i, a.id))
    if skip_connection:
        try_skip_connection(TPG, collide, connect)
    return TPG


def get_causal_actions(action: Action, TPG, for_ancestors: bool) -> Set[Tuple[int, str]]:
    """smoother.jl:220-242"""
    tables: List[Dict[str, Set[Tuple[int, str]]]] = [dict() for _ in TPG]

    def f(i, id_):
        a = _find(TPG, i, id_)
        if id_ in tables[i]:
            return tables[i][id_]
        causal: Set[Tuple[int, str]] = set()
        for (j, id_j) in (a.predecessors if for_ancestors else a.successors):
            causal.add((j, id_j))
            causal |= f(j, id_j)
        tables[i][id_] = causal
        return causal

    return f(action.agent, action.id)


def try_skip_connection(TPG, collide, connect) -> None:
    """smoother.jl:135-217.  Deviation: an agent without any action has no `TPG[j][end]`
    (the reference would raise a BoundsError at :186); it is skipped here."""
    N = len(TPG)
    for i in range(N):
        k = 0                                            # Julia k = 1; a1 = TPG[i][k-1], a2 = TPG[i][k]
        while k + 1 < len(TPG[i]):
            k += 1
            a1, a2 = TPG[i][k - 1], TPG[i][k]
            if not connect(a1.frm.q, a2.to.q, i):                          # :150
                continue
            if any(j != i for (j, _) in a2.predecessors):                  # :153
                continue
            if any(j != i for (j, _) in a1.successors):                    # :154
                continue
            a3 = Action(get_action_id(a1.frm, a2.to, a1.t), a1.frm, a2.to, a1.t, a1.agent,
                        a1.predecessors, a2.successors)
            causal = get_causal_actions(a1, TPG, True) | get_causal_actions(a2, TPG, False)
            causal = {c for c in causal if c[0] != i}
            conflicted = False
            for j in range(N):                                             # :174-191
                if j == i:
                    continue
                for a4 in [a for a in TPG[j] if (j, a.id) not in causal]:
                    if collide(a3.frm.q, a3.to.q, a4.frm.q, a4.to.q, i, j):
                        conflicted = True
                        break
                if conflicted:
                    continue
                if TPG[j]:
                    last = TPG[j][-1].to.q
                    if collide(a3.frm.q, a3.to.q, last, last, i, j):
                        conflicted = True
                        break
            if conflicted:
                continue
            TPG[i][k - 1:k + 1] = [a3]                                     # :197-198
            for (j, id_) in a3.predecessors:                               # :201-212
                a4 = _find(TPG, j, id_)
                a4.successors.remove((i, a1.id))
                a4.successors.append((i, a3.id))
            for (j, id_) in a3.successors:
                a4 = _find(TPG, j, id_)
                a4.predecessors.remove((i, a2.id))
                a4.predecessors.append((i, a3.id))
            k -= 1


def get_greedy_solution(TPG, config_goal) -> List[List[Vertex]]:
    """smoother.jl:258-295"""
    N = len(TPG)
    idx = [0] * N
    solution = [[(TPG[i][0].frm if TPG[i] else Vertex(np.asarray(config_goal[i], float), -1))
                 for i in range(N)]]
    while not all(idx[i] >= len(TPG[i]) for i in range(N)):
        config, last = [], list(idx)
        for i in range(N):
            if TPG[i]:
                a = TPG[i][min(idx[i], len(TPG[i]) - 1)]
                if all([b.id for b in TPG[j]].index(id_) < last[j] for (j, id_) in a.predecessors):
                    idx[i] = min(idx[i] + 1, len(TPG[i]))
                    config.append(a.to)
                else:
                    config.append(a.frm)
            else:
                config.append(Vertex(np.asarray(config_goal[i], float), -1))
        solution.append(config)
    return solution


def get_tpg_cost(TPG, dist) -> Optional[Dict[str, float]]:
    """smoother.jl:339-363 (dynamic programming over the dependencies)"""
    if TPG is None:
        return None
    tables: List[Dict[str, float]] = [dict() for _ in TPG]

    def f(i, id_):
        a = _find(TPG, i, id_)
        c = dist(a.frm.q, a.to.q)
        if a.predecessors:
            c += max(tables[j][id_j] if id_j in tables[j] else f(j, id_j) for (j, id_j) in a.predecessors)
        tables[i][id_] = c
        return c

    arr = [f(i, TPG[i][-1].id) if TPG[i] else 0.0 for i in range(len(TPG))]
    s = 0.0
    for x in arr:
        s += x
    return {"sum_of_cost": s, "makespan": max(arr)}


def smoothing(solution, connect, collide, dist):
    """smoother.jl:379-411.  The loop returns in its first pass: sum_of_cost_last starts at Inf and
    the exit test is `sum_of_cost_last >= cost[:sum_of_cost]` (:399)."""
    if solution is None:
        return None
    config_goal = [v.q for v in solution[-1]]
    TPG = get_temporal_plan_graph(solution, collide, connect)
    sol = get_greedy_solution(TPG, config_goal)
    return TPG, sol, get_tpg_cost(TPG, dist)


def oracle_closures(orc):
    """collide / connect / dist closures of the reference's signatures over an oracle.Oracle"""
    collide = lambda a, b, c, d, i, j: bool(orc.collide_pair(a, b, c, d, i, j))
    connect = lambda a, b, i: bool(orc.connect_edge(a, b, i))
    dist = lambda a, b: float(orc.state_dist(a, b))
    return collide, connect, dist
