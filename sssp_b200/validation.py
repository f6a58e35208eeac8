"""validate / is_valid_instance of the reference (src/utils.jl:261-367) on the batched GPU path.

Same signatures and return values as the reference; `validate` additionally exposes the first
failing check (`validate.last_failure`) that the reference only prints as a warning."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import lib as L
from .closures import Context, Node, _states


def validate(config_init: Sequence, connect, collide, check_goal, solution) -> bool:
    """utils.jl:320-367.  `solution` = Vector{Vector{Node}} (one configuration per step) or None."""
    validate.last_failure = None
    if solution is None:
        return True
    ctx: Context = connect.ctx
    N, T, d = len(config_init), len(solution), ctx.d
    sol = np.stack([_states(Q, d) for Q in solution])            # [T][N][d]
    if not np.array_equal(sol[0], _states(config_init, d)):        # :333-336
        validate.last_failure = ("initial configuration",)
        return False
    if not check_goal(solution[-1]):                               # :339-343
        validate.last_failure = ("last configuration",)
        return False
    if T < 2:
        return True
    # Transition checks (:345-363), each through the closure that owns it: the connect tests run
    # on connect's context (obstacles, max_dist), the collide tests on COLLIDE's context -- a
    # collide closure built with its own step_dist / safety_dist / base positions decides, exactly
    # as in the reference where the two closures are independent (arm22.jl:29-37 vs :93-99).
    ag = np.tile(np.arange(N, dtype=np.int32), T - 1)
    con = connect.ctx.connect_edges(ag, sol[:-1].reshape(-1, d), sol[1:].reshape(-1, d)).reshape(T - 1, N)
    hit, first = collide.ctx.collide_configs(sol[:-1], sol[1:])
    for t in range(1, T):                      # the reference's order: connect (i ascending), then collide
        bad = np.flatnonzero(con[t - 1] == 0)
        if bad.size:
            validate.last_failure = ("not connected", t, int(bad[0]))
            return False
        if hit[t - 1]:
            fp = int(first[t - 1])
            validate.last_failure = ("colliding", t, fp // N, fp % N)
            return False
    return True


validate.last_failure = None


def is_valid_instance(config_init: Sequence, config_goal: Sequence, connect, collide) -> bool:
    """utils.jl:261-307 on closures built by gen_connect / gen_collide: starts and goals are
    free (connect(q, q, i)) and pairwise collision-free (collide(q_i, q_i, q_j, q_j, i, j))."""
    ctx: Context = connect.ctx
    N, d = len(config_init), ctx.d
    init, goal = _states(config_init, d), _states(config_goal, d)
    ag = np.tile(np.arange(N, dtype=np.int32), 2)
    q = np.concatenate([init, goal])
    if not ctx.connect_edges(ag, q, q).all():
        return False
    Q = np.stack([init, goal])
    hit, _ = collide.ctx.collide_configs(Q, Q)
    return not hit.any()
