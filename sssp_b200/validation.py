"""validate / is_valid_instance of the reference (src/utils.jl:261-367) on the batched GPU path.

Same signatures and return values as the reference; `validate` additionally exposes the first
failing check (`validate.last_failure`) that the reference only prints as a warning."""
from __future__ import annotations

import ctypes as C
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
    ok, kind, i, j = C.c_int32(1), C.c_int32(0), C.c_int32(-1), C.c_int32(-1)
    t = C.c_int64(-1)
    L.check(ctx._lib.mrmp_validate_solution(ctx._h, T, L.ptr(sol, L._dp), C.byref(ok), C.byref(t),
                                            C.byref(kind), C.byref(i), C.byref(j)))
    if not ok.value:
        validate.last_failure = ("not connected", int(t.value), int(i.value)) if kind.value == 1 \
            else ("colliding", int(t.value), int(i.value), int(j.value))
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
