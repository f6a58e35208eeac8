"""Literal pure-Python transliteration of the reference's Julia source for the hot path.

CPU ORACLE (test infrastructure, NOT product code).  Second, independent restatement used to
cross-check oracle/mrmp_oracle.c: it keeps the reference's vector-valued formulation
(`e * a + (1 - e) * b - c` on whole vectors, comprehension + any/minimum) instead of the C
oracle's scalarised loops.  Python floats are IEEE binary64 and CPython never fuses a*b+c,
so both must agree bit for bit.  Pure-Python loops: small cases only.

Each function cites the reference file:line (paths under /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import List, Sequence

Vec = List[float]


# --- Julia Base / LinearAlgebra pieces (not in /root/reference) -------------
def norm(v: Sequence[float]) -> float:
    """LinearAlgebra.generic_norm2 for short vectors."""
    maxabs = 0.0
    for x in v:
        a = abs(x)
        maxabs = math.nan if (math.isnan(a) or math.isnan(maxabs)) else max(maxabs, a)
    if maxabs == 0.0 or math.isinf(maxabs):
        return maxabs
    if math.isfinite(len(v) * maxabs * maxabs) and maxabs * maxabs != 0.0:
        s = v[0] * v[0]
        for x in v[1:]:
            s += x * x
        return math.sqrt(s)
    t = abs(v[0]) / maxabs
    s = t * t
    for x in v[1:]:
        t = abs(x) / maxabs
        s += t * t
    return maxabs * math.sqrt(s)


def dot(a: Sequence[float], b: Sequence[float]) -> float:
    s = a[0] * b[0]
    for x, y in zip(a[1:], b[1:]):
        s += x * y
    return s


def vsub(a, b) -> Vec:
    return [x - y for x, y in zip(a, b)]


def vadd(a, b) -> Vec:
    return [x + y for x, y in zip(a, b)]


def smul(s, a) -> Vec:
    return [s * x for x in a]


def jl_minimum(xs: Sequence[float]) -> float:
    if any(math.isnan(x) for x in xs):
        return math.nan
    return min(xs)


# --- src/utils.jl:17-94 -------------------------------------------------------------------
def dist_pp(a, b) -> float:  # :17-19
    return norm(vsub(a, b))


def dist_sp(a, b, c) -> float:  # :44-57
    df_0 = dot(vsub(a, b), vsub(b, c))
    df_1 = dot(vsub(a, b), vsub(a, c))
    e = 1
    if df_0 * df_1 < 0:
        e = -dot(vsub(a, b), vsub(b, c)) / dot(vsub(a, b), vsub(a, b))
    elif df_0 >= 0:
        e = 0
    return norm(vsub(vadd(smul(e, a), smul(1 - e, b)), c))


def dist_ss(p1_1, p1_2, p2_1, p2_2) -> float:  # :59-94
    v1 = vsub(p1_2, p1_1)
    v2 = vsub(p2_2, p2_1)
    D1 = dot(vsub(p2_1, p1_1), v1)
    D2 = dot(vsub(p2_1, p1_1), v2)
    Dv = dot(v1, v2)
    V1 = dot(v1, v1)
    V2 = dot(v2, v2)
    D = V1 * V2 - Dv * Dv
    if D > 0:
        t1 = (D1 * V2 - D2 * Dv) / D
        t2 = (D1 * Dv - D2 * V1) / D
        if 0 <= t1 <= 1 and 0 <= t2 <= 1:
            Q1 = vadd(p1_1, smul(t1, v1))
            Q2 = vadd(p2_1, smul(t2, v2))
            return norm(vsub(Q1, Q2))
    return jl_minimum([
        dist_sp(p1_1, p1_2, p2_1),
        dist_sp(p1_1, p1_2, p2_2),
        dist_sp(p2_1, p2_2, p1_1),
        dist_sp(p2_1, p2_2, p1_2),
    ])


# --- point models: point2d.jl:42-69, point3d.jl:47-80, point.jl:5-36 -----------------------
class PointInstance:
    def __init__(self, rads, obstacles):
        self.rads = list(rads)
        self.obstacles = [list(o) for o in obstacles]  # x, y, (z), r

    def connect_point(self, q, i) -> bool:  # point2d.jl:49-57
        r = self.rads[i]
        if any((x < r or 1 - r < x) for x in q):
            return False
        if any([dist_pp(list(q), o[:-1]) < o[-1] + r for o in self.obstacles]):
            return False
        return True

    def connect_edge(self, q_from, q_to, i) -> bool:  # point2d.jl:59-66
        r = self.rads[i]
        if any((x < r or 1 - r < x) for x in q_to):
            return False
        if any([dist_sp(list(q_from), list(q_to), o[:-1]) < o[-1] + r for o in self.obstacles]):
            return False
        return True

    def collide_pair(self, a, b, c, d, i, j) -> bool:  # point.jl:9-12
        return dist_ss(list(a), list(b), list(c), list(d)) < self.rads[i] + self.rads[j]

    def collide_static(self, qi, qj, i, j) -> bool:  # point.jl:14-16
        return dist_pp(list(qi), list(qj)) < self.rads[i] + self.rads[j]

    def collide_configs(self, Qf, Qt) -> bool:  # point.jl:18-25
        N = len(self.rads)
        for i in range(N):
            for j in range(i + 1, N):
                if self.collide_pair(Qf[i], Qt[i], Qf[j], Qt[j], i, j):
                    return True
        return False

    def collide_config_move(self, Q, q_to, i) -> bool:  # point.jl:27-33
        q_from = Q[i]
        for j in range(len(self.rads)):
            if j != i and dist_sp(list(q_from), list(q_to), list(Q[j])) < self.rads[i] + self.rads[j]:
                return True
        return False

    def prm_neighbors(self, nodes, rad, i):  # prm.jl:202-210
        nv = len(nodes)
        nb = [[] for _ in range(nv)]
        for j in range(nv):
            for k in range(nv):
                if j == k:
                    continue
                if (rad is None or dist_pp(list(nodes[j]), list(nodes[k])) <= rad) and \
                        self.connect_edge(nodes[j], nodes[k], i):
                    nb[j].append(k)
        return nb
