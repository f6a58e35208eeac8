"""Solution smoother of the reference (src/smoother.jl) on the batched GPU path.

`smoothing(solution, connect, collide)` builds the temporal plan graph (TPG), tries to merge
consecutive actions, samples a greedy solution from the graph and returns (TPG, solution, cost)
like smoother.jl:379-411.  The reference issues one `collide` call per (action, later action of
another agent) pair until the first hit (smoother.jl:78-106) and, per merge candidate, one call
per remaining action of every other agent (:174-191).  Here

  * the whole type-2 dependency scan is ONE launch: mrmp_collide_cross_reduce in FIRST mode,
    rows = columns = all actions, group = the other agent, keys = time steps
    (first later action of agent j that collides with action_self, for every pair at once);
  * every merge candidate is one cross-product launch (the candidate against all actions and last
    locations of the other agents); the causal-set filter is applied to the returned bits.

The graph surgery stays on the host exactly as written in the reference.

Conventions: agents, vertex ids and time steps are 0-based (the reference is 1-based);
`get_instructions` converts to the reference's numbering at the JSON boundary.  Where Julia
iterates `values(Dict)` (hash order, smoother.jl:119) predecessors are listed by ascending agent.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from .closures import Context, Node, _states


@dataclass
class Action:          # smoother.jl:13-21
    id: str
    frm: Node          # `from` in the reference
    to: Node
    t: int
    agent: int
    predecessors: List[Tuple[int, str]] = field(default_factory=list)
    successors: List[Tuple[int, str]] = field(default_factory=list)


def get_action_id(frm: Node, to: Node, t: int) -> str:   # smoother.jl:23-29
    return "%d_%d_%d" % (frm.id, to.id, t)


def _find(TPG, j, id_) -> Action:
    for a in TPG[j]:
        if a.id == id_:
            return a
    raise KeyError((j, id_))


def get_temporal_plan_graph(solution: Sequence[Sequence[Node]], collide, connect,
                            skip_connection: bool = True) -> List[List[Action]]:
    """smoother.jl:32-132 with the type-2 scan batched into one launch."""
    ctx: Context = collide.ctx
    N = len(solution[0])
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
    # ---- type-2 dependency (:78-106): first colliding LATER action of every other agent
    flat = [a for acts in TPG for a in acts]
    if flat:
        agent = np.array([a.agent for a in flat], np.int32)
        key = np.array([a.t for a in flat], np.int32)
        qf = _states([a.frm.q for a in flat], ctx.d)
        qt = _states([a.to.q for a in flat], ctx.d)
        first = ctx.collide_cross_reduce(agent, qf, qt, agent, qf, qt, "first", col_group=agent,
                                         n_groups=N, row_key=key, col_key=key)
        r = 0
        for i in range(N):
            for a_self in TPG[i]:
                for j in range(N):
                    c = int(first[r, j])
                    if j != i and c >= 0:
                        flat[c].predecessors.append((a_self.agent, a_self.id))
                r += 1
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


def get_ancestors(action, TPG):      # smoother.jl:244-249
    return get_causal_actions(action, TPG, True)


def get_descendants(action, TPG):    # smoother.jl:251-256
    return get_causal_actions(action, TPG, False)


def try_skip_connection(TPG: List[List[Action]], collide, connect) -> None:
    """smoother.jl:135-217.  The collision test of one merge candidate (:174-191: every action of
    every other agent outside the causal set, plus each agent's last location) is one
    cross-product launch; its verdict `conflicted` is an OR, so the reference's early exits do not
    change it.  An agent without actions has no `TPG[j][end]` (the reference would raise a
    BoundsError at :186): it contributes no last-location test."""
    ctx: Context = collide.ctx
    N = len(TPG)
    for i in range(N):
        k = 0
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
            causal = get_ancestors(a1, TPG) | get_descendants(a2, TPG)
            others = [(j, a.frm.q, a.to.q) for j in range(N) if j != i
                      for a in TPG[j] if (j, a.id) not in causal]
            others += [(j, TPG[j][-1].to.q, TPG[j][-1].to.q) for j in range(N) if j != i and TPG[j]]
            conflicted = False
            if others:
                bits = ctx.collide_cross([i], [a3.frm.q], [a3.to.q], [o[0] for o in others],
                                         [o[1] for o in others], [o[2] for o in others], 0)
                conflicted = bool(bits.any())
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


def get_greedy_solution(TPG: List[List[Action]], config_goal) -> List[List[Node]]:
    """smoother.jl:258-295"""
    N = len(TPG)
    idx = [0] * N
    solution = [[(TPG[i][0].frm if TPG[i] else Node(config_goal[i], -1, [])) for i in range(N)]]
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
                config.append(Node(config_goal[i], -1, []))
        solution.append(config)
    return solution


def get_tpg_cost(TPG: Optional[List[List[Action]]], ctx: Context) -> Optional[Dict[str, float]]:
    """smoother.jl:339-363; the model's dist of every action in one batched call"""
    if TPG is None:
        return None
    flat = [a for acts in TPG for a in acts]
    d = ctx.state_dist([a.frm.q for a in flat], [a.to.q for a in flat]) if flat else []
    length = {(a.agent, a.id): float(x) for a, x in zip(flat, d)}
    tables: List[Dict[str, float]] = [dict() for _ in TPG]

    def f(i, id_):
        a = _find(TPG, i, id_)
        c = length[(i, id_)]
        if a.predecessors:
            c += max(tables[j][id_j] if id_j in tables[j] else f(j, id_j) for (j, id_j) in a.predecessors)
        tables[i][id_] = c
        return c

    arr = [f(i, TPG[i][-1].id) if TPG[i] else 0.0 for i in range(len(TPG))]
    s = 0.0
    for x in arr:
        s += x
    return {"sum_of_cost": s, "makespan": max(arr) if arr else 0.0}


def smoothing(solution, connect, collide):
    """smoother.jl:379-411 -> (TPG, solution, cost).  The reference's loop returns in its first
    pass (sum_of_cost_last starts at Inf and the exit test is `>=`, :399)."""
    if solution is None:
        return None
    config_goal = [v.q for v in solution[-1]]
    TPG = get_temporal_plan_graph(solution, collide, connect)
    return TPG, get_greedy_solution(TPG, config_goal), get_tpg_cost(TPG, collide.ctx)


def get_instructions(TPG: List[List[Action]], field_x: float = 0.0, field_y: float = 0.0,
                     field_size: float = 1.0) -> List[List[dict]]:
    """scripts/server.jl:39-55: the `instructions` array of the planning server's reply for
    StatePoint2D agents, in the reference's numbering (agents, vertex ids and time steps of the
    action ids are 1-based there) and rescaled to the field."""
    def ref_id(a_id: str) -> str:
        f, t, s = a_id.split("_")
        return "%d_%d_%d" % (int(f) + 1, int(t) + 1, int(s) + 1)

    def one(a: Action) -> dict:
        return {"id": ref_id(a.id),
                "x_from": a.frm.q[0] * field_size + field_x, "y_from": a.frm.q[1] * field_size + field_y,
                "x_to": a.to.q[0] * field_size + field_x, "y_to": a.to.q[1] * field_size + field_y,
                "pre": [[j + 1, ref_id(i)] for (j, i) in a.predecessors],
                "suc": [[j + 1, ref_id(i)] for (j, i) in a.successors]}

    return [[one(a) for a in acts] for acts in TPG]
