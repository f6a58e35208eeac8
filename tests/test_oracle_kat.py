"""Pin the oracle: the reference's only known-answer vector (test/dist.jl:4-9) plus KATs
derived by hand from src/utils.jl:44-94 and src/models/point2d.jl:49-66 (these
are labelled DERIVED, not reference vectors)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as R


def test_reference_kat_dist_jl():
    # /root/reference/test/dist.jl:4-9 : dist(p1_1, p1_2, p2_1, p2_2) == 0
    assert O.dist_ss([-1.0, -2.0], [1.0, 2.0], [1.0, -2.0], [-1.0, 2.0]) == 0
    assert R.dist_ss([-1.0, -2.0], [1.0, 2.0], [1.0, -2.0], [-1.0, 2.0]) == 0


@pytest.mark.parametrize("a,b,c,want", [
    ([0, 0], [1, 0], [0.5, 1], 1.0),    # interior, e = 0.5
    ([0, 0], [1, 0], [2, 0], 1.0),      # e = 0 -> endpoint b
    ([0, 0], [1, 0], [-1, 0], 1.0),     # e = 1 -> endpoint a
    ([0, 0], [0, 0], [3, 4], 5.0),      # degenerate a == b
    ([0, 0, 0], [0, 0, 2], [1, 0, 1], 1.0),
])
def test_derived_segment_point(a, b, c, want):
    assert O.dist_sp(a, b, c) == want
    assert R.dist_sp(list(map(float, a)), list(map(float, b)), list(map(float, c))) == want


def test_derived_segment_segment_fallbacks():
    # parallel: D == 0 -> endpoint fallback
    assert O.dist_ss([0, 0], [1, 0], [0, 1], [1, 1]) == 1.0
    # stationary second segment: V2 == 0 -> fallback
    assert O.dist_ss([0, 0], [1, 0], [0.5, 2], [0.5, 2]) == 2.0
    # both degenerate
    assert O.dist_ss([0, 0], [0, 0], [3, 4], [3, 4]) == 5.0


def test_derived_point2d_connect():
    o = O.Oracle(O.POINT2D, [0.1], [[0.5, 0.5, 0.1]])
    assert not o.connect_edge([0.2, 0.5], [0.8, 0.5], 0)      # through the obstacle
    assert o.connect_edge([0.2, 0.2], [0.8, 0.2], 0)          # distance 0.3 >= 0.2
    assert not o.connect_edge([0.2, 0.2], [0.95, 0.2], 0)     # q_to out of field
    assert o.connect_edge([0.95, 0.2], [0.8, 0.2], 0)         # q_from is NOT bounds-checked
    assert o.connect_point([0.5, 0.2], 0)
    assert not o.connect_point([0.05, 0.5], 0)
    assert not o.connect_point([0.5, 0.65], 0)                # inside o.r + rad


def test_strictness_at_threshold():
    # distance exactly equal to the threshold is NOT a collision (strict <, point.jl:11,15)
    o = O.Oracle(O.POINT2D, [0.25, 0.25], np.zeros((0, 3)))
    assert not o.collide_static([0.0, 0.0], [0.5, 0.0], 0, 1)
    assert o.collide_static([0.0, 0.0], [math.nextafter(0.5, 0), 0.0], 0, 1)
    assert not o.collide_pair([0, 0], [0, 1], [0.5, 0], [0.5, 1], 0, 1)
    assert o.collide_pair([0, 0], [0, 1], [math.nextafter(0.5, 0), 0], [0.5, 1], 0, 1)


def test_collide_configs_first_pair_order():
    # three agents on a line, all overlapping: first hit is (0,1) in (i asc, j asc) order
    o = O.Oracle(O.POINT2D, [0.1, 0.1, 0.1], np.zeros((0, 3)))
    Q = [[0.5, 0.5], [0.55, 0.5], [0.6, 0.5]]
    hit, fp = o.collide_configs(Q, Q)
    assert hit and fp == 0 * 3 + 1
    Q2 = [[0.1, 0.1], [0.5, 0.5], [0.55, 0.5]]
    hit, fp = o.collide_configs(Q2, Q2)
    assert hit and fp == 1 * 3 + 2
    Q3 = [[0.1, 0.1], [0.5, 0.5], [0.9, 0.9]]
    assert o.collide_configs(Q3, Q3) == (False, -1)


def test_norm_underflow_path():
    # generic_norm2 rescales when the plain sum of squares underflows
    v = [3e-200, 4e-200]
    assert O.norm(v) == pytest.approx(5e-200, rel=1e-15)
    assert O.norm(v) == R.norm(v)
    assert O.norm([0.0, 0.0]) == 0.0


def test_arm22_static_equals_edge_with_same_endpoints():
    # D == 0 -> 0/0 = NaN step is a silent no-op, only e = 1 is effective
    rng = np.random.default_rng(3)
    rads = [0.12, 0.1]
    base = [[0.4, 0.5], [0.65, 0.5]]
    obs = [[0.5, 0.8, 0.08]]
    o = O.Oracle(O.ARM22, rads, obs, base=base)
    for _ in range(200):
        q = rng.uniform(0, 2 * math.pi, 2)
        i = int(rng.integers(0, 2))
        assert o.connect_edge(q, q, i) == o.connect_point(q, i)
    s = O.step_schedule(0.01, 0.0)
    assert len(s) == 2 and math.isnan(s[0]) and s[1] == 1.0
