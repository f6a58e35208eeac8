"""The oracle's msun-style sin/cos/atan2 (oracle/orc_trig.c) against the host libm."""
import math

import numpy as np

from oracle import oracle as O


def _ulps(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / math.ulp(b)


def test_sin_cos_within_1ulp():
    L = O.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-4 * math.pi, 4 * math.pi, 60000),
                         rng.uniform(-1e-4, 1e-4, 2000), rng.uniform(-1e5, 1e5, 5000),
                         np.arange(-16, 17) * (math.pi / 4), [0.0, -0.0, 1e-300]])
    worst = 0.0
    for x in xs:
        worst = max(worst, _ulps(L.orc_sin(x), math.sin(x)), _ulps(L.orc_cos(x), math.cos(x)))
    assert worst <= 1.0


def test_atan2_within_1ulp_and_quadrants():
    L = O.lib()
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(40000):
        t = rng.uniform(-math.pi, math.pi)
        y, x = math.sin(t), math.cos(t)
        worst = max(worst, _ulps(L.orc_atan2(y, x), math.atan2(y, x)))
    assert worst <= 1.0
    assert L.orc_atan2(0.0, -1.0) == math.pi
    assert L.orc_atan2(-0.0, -1.0) == -math.pi
    assert L.orc_atan2(1.0, 0.0) == math.pi / 2
    assert L.orc_atan2(0.0, 1.0) == 0.0


def test_diff_angles_range_and_antisymmetry():
    rng = np.random.default_rng(2)
    for _ in range(20000):
        a, b = rng.uniform(0, 2 * math.pi, 2)
        dab = O.diff_angles(a, b)
        assert -math.pi <= dab <= math.pi
        assert dab == -O.diff_angles(b, a) or abs(abs(dab) - math.pi) < 1e-12


def test_step_schedule_rules():
    # literal fallback: len = round(D/step)+1, minus one on overshoot; elements k*step/D
    s = O.step_schedule(0.01, 0.05)   # rational path: 0.05 == 1/20
    assert np.allclose(s[:-1], [0, .2, .4, .6, .8, 1.0]) and s[-1] == 1.0
    D = 0.123456789
    s = O.step_schedule(0.01, D)
    assert len(s) == 13 + 1 and s[3] == (3 * 0.01) / D and s[-1] == 1.0
    # rational lift for short-rational D: elements are RN(k/100), e.g. k = 35
    s = O.step_schedule(0.01, 0.5)
    assert len(s) == 51 + 1 and s[35] == 0.35 / 0.5 and (35 * 0.01) != 0.35
    # the worst case D = sqrt(2) needs 143 entries
    assert len(O.step_schedule(0.01, math.sqrt(2.0))) == 143
