"""The device-side arithmetic headers (geom.cuh, trig.cuh, arm_model.cuh), compiled for the
HOST by tests/hostprobe/probe.cu, must agree bit for bit with the oracle.  This runs in the
GPU-less container and pins the source the kernels are built from; the `-m gpu` tests then
check the same functions as compiled for sm_100a."""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import make_arm_instance, near_threshold_pairs
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostprobe", "probe.cu")
OUT = os.path.join(HERE, "_build", "libhostprobe.so")
_dp = C.POINTER(C.c_double)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


class ProbeArm(C.Structure):
    _fields_ = [("n_obs", C.c_int), ("ox", _dp), ("oy", _dp), ("orad", _dp), ("t2", _dp),
                ("step", C.c_double), ("safety", C.c_double), ("safety_t2", C.c_double),
                ("max_dist", C.c_double)]


@pytest.fixture(scope="module")
def probe():
    nvcc = _nvcc()
    if nvcc is None:
        pytest.skip("nvcc not available")
    csrc = os.path.join(HERE, "..", "sssp_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(p) > os.path.getmtime(OUT) for p in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                               "-O2", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o",
                               OUT, SRC])
    lib = C.CDLL(OUT)
    lib.probe_sincos.argtypes = [C.c_double, _dp, _dp]
    lib.probe_atan2.restype = C.c_double
    lib.probe_atan2.argtypes = [C.c_double, C.c_double]
    lib.probe_diff_angles.restype = C.c_double
    lib.probe_diff_angles.argtypes = [C.c_double, C.c_double]
    lib.probe_segseg_lt.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double]
    lib.probe_segseg_lt_sqrt.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_double]
    lib.probe_segpt_lt.argtypes = [_dp, _dp, _dp, C.c_int, C.c_double, C.c_double]
    vp = C.c_void_p
    lib.probe_segseg_batch.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp]
    lib.probe_segpt_batch.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp]
    lib.probe_segseg_filter_batch.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.probe_bound32_batch.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.probe_step_schedule.restype = C.c_int64
    lib.probe_step_schedule.argtypes = [C.c_double, C.c_double, _dp, C.c_int64]
    P = C.POINTER(ProbeArm)
    lib.probe_arm_connect_point.argtypes = [P, _dp] + [C.c_double] * 3
    lib.probe_arm_connect_edge.argtypes = [P, _dp, _dp] + [C.c_double] * 3
    lib.probe_arm_collide_static.argtypes = [P, _dp, _dp] + [C.c_double] * 6
    lib.probe_arm_collide_pair.argtypes = [P, _dp, _dp, _dp, _dp] + [C.c_double] * 6
    lib.probe_chain.argtypes = [C.c_int, C.c_void_p, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.probe_arm_dist.restype = C.c_double
    lib.probe_arm_dist.argtypes = [_dp, _dp]
    lib.probe_arm_mid_status.argtypes = [_dp, _dp, _dp]
    return lib


def p(a):
    return a.ctypes.data_as(_dp)


def sq_lt_threshold(thr):
    """min{ s : RN(sqrt(s)) >= thr } (abi.cu: sq_lt_threshold)."""
    if not thr > 0.0:
        return 0.0
    s = thr * thr
    while s > 0.0 and math.sqrt(s) >= thr:
        s = math.nextafter(s, -math.inf)
    while math.sqrt(s) < thr:
        s = math.nextafter(s, math.inf)
    return s


def bits(x):
    return np.float64(x).view(np.uint64)


def test_trig_bit_identical_to_oracle(probe):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-20, 20, 100000), rng.uniform(-1e-3, 1e-3, 5000),
                         rng.uniform(-1e5, 1e5, 20000),
                         np.array([0.0, -0.0, 1e-30, -1e-30, math.pi, -math.pi, math.pi / 2,
                                   math.pi / 4, 0.7853981633974483, 0.7853981633974484,
                                   2 * math.pi, 1e-9, 2.0 ** -26, 2.0 ** -27])])
    L = O.lib()
    s, c = C.c_double(), C.c_double()
    for x in xs:
        probe.probe_sincos(x, C.byref(s), C.byref(c))
        assert bits(s.value) == bits(L.orc_sin(x)), x
        assert bits(c.value) == bits(L.orc_cos(x)), x
    ys = rng.normal(size=60000) * rng.choice([1e-12, 1e-3, 1.0, 1e3], 60000)
    zs = rng.normal(size=60000) * rng.choice([1e-12, 1e-3, 1.0, 1e3], 60000)
    ys[:8] = [0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1.0, 3.0]
    zs[:8] = [1.0, 1.0, -1.0, -1.0, 0.0, 0.0, 1.0, -0.0]
    for y, z in zip(ys, zs):
        assert bits(probe.probe_atan2(y, z)) == bits(L.orc_atan2(y, z)), (y, z)
    a = rng.uniform(0, 2 * math.pi, 40000)
    b = rng.uniform(0, 2 * math.pi, 40000)
    a[:3], b[:3] = [1.0, math.pi, 0.0], [1.0, 0.0, math.pi]
    for t1, t2 in zip(a, b):
        assert bits(probe.probe_diff_angles(t1, t2)) == bits(L.orc_diff_angles(t1, t2))


def test_step_schedule_matches_oracle(probe):
    rng = np.random.default_rng(6)
    Ds = list(rng.uniform(0, math.sqrt(2), 20000)) + [k / 100 for k in range(0, 143)] + \
        [0.0, 0.29, 0.58, 0.5, 1.0, 1e-9, 0.01, 0.005, 0.015, 1.4142135623730951, 1 / 3, 2 / 7]
    out = np.empty(8192)
    for step in (0.01, 0.05, 0.013, 1 / 128):
        for D in Ds:
            n = probe.probe_step_schedule(step, D, p(out), out.size)
            ref = O.step_schedule(step, D)
            assert n == ref.size, (step, D)
            assert np.array_equal(out[:n].view(np.uint64), ref.view(np.uint64)), (step, D)


def test_segment_tests_match_oracle(probe):
    rng = np.random.default_rng(7)
    n = 20000
    for d in (2, 3):
        rads = rng.uniform(0.015, 0.03, 8)
        ai, aj = rng.integers(0, 8, n), rng.integers(0, 8, n)
        a, b, c, e = near_threshold_pairs(rng, n, rads, ai, aj, d)
        a[: n // 2] = rng.random((n // 2, d))        # half plain random
        b[: n // 2] = a[: n // 2] + rng.normal(size=(n // 2, d)) * 0.05
        b[:50] = a[:50]                               # degenerate first segment
        e[50:100] = c[50:100]                         # stationary second segment
        for k in range(n):
            thr = rads[ai[k]] + rads[aj[k]]
            want = O.dist_ss(a[k], b[k], c[k], e[k]) < thr
            got = probe.probe_segseg_lt(p(a[k]), p(b[k]), p(c[k]), p(e[k]), d, thr,
                                        sq_lt_threshold(thr))
            got2 = probe.probe_segseg_lt_sqrt(p(a[k]), p(b[k]), p(c[k]), p(e[k]), d, thr)
            assert bool(got) == want and bool(got2) == want, (d, k)
            want = O.dist_sp(a[k], b[k], c[k]) < thr
            got = probe.probe_segpt_lt(p(a[k]), p(b[k]), p(c[k]), d, thr, sq_lt_threshold(thr))
            assert bool(got) == want, (d, k)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("shift", [0.0, 1.0e3, 1.0e5])
def test_quick_tests_never_change_a_verdict(probe, d, shift):
    """The division-free quick tests (geom.cuh) against the oracle on 300k pairs per case:
    radii from 1e-3 (quick path just enabled) to 0.2, contact offsets straddling the 1e-6
    hand-over band, and the whole scene translated by up to 1e5 (the reference's own rounding
    error grows with the coordinate magnitude; the band must still cover it)."""
    from helpers import near_threshold_pairs
    rng = np.random.default_rng(11 + d)
    N, n = 12, 300_000
    rads = np.concatenate([[5e-4, 5.1e-4, 1e-3], rng.uniform(0.002, 0.1, N - 3)])
    ai = rng.integers(0, N, n).astype(np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    a, b, c, e = near_threshold_pairs(rng, n, rads, ai, aj, d)
    h = n // 3
    a[:h] = rng.random((h, d))
    b[:h] = a[:h] + rng.normal(size=(h, d)) * 0.05
    c[:h] = a[:h] + rng.normal(size=(h, d)) * 0.05
    e[:h] = c[:h] + rng.normal(size=(h, d)) * 0.05
    a, b, c, e = (np.ascontiguousarray(x + shift) for x in (a, b, c, e))
    thr = (rads[:, None] + rads[None, :]).copy()
    T2 = np.array([[sq_lt_threshold(t) for t in row] for row in thr])
    orc = O.Oracle(O.POINT2D if d == 2 else O.POINT3D, rads, np.zeros((0, d + 1)))
    want = orc.collide_pairs(ai, aj, a, b, c, e, nthreads=0)
    got = np.empty(n, np.uint8)
    probe.probe_segseg_batch(n, d, a.ctypes.data, b.ctypes.data, c.ctypes.data, e.ctypes.data,
                             ai.ctypes.data, aj.ctypes.data, N, thr.ctypes.data, T2.ctypes.data,
                             got.ctypes.data)
    assert np.array_equal(got, want) and 0.05 * n < want.sum() < 0.95 * n
    # segment-point form: collide(Q, q_to, i) of one candidate against one static agent
    Q = np.zeros((N, d))
    wantp = np.array([O.dist_sp(a[k], b[k], c[k]) < thr[ai[k], aj[k]] for k in range(0, n, 10)], np.uint8)
    gotp = np.empty(n, np.uint8)  # 255 = branchy and branch-free device forms disagree
    probe.probe_segpt_batch(n, d, a.ctypes.data, b.ctypes.data, c.ctypes.data, ai.ctypes.data,
                            aj.ctypes.data, N, thr.ctypes.data, T2.ctypes.data, gotp.ctypes.data)
    assert gotp.max() <= 1
    assert np.array_equal(gotp[::10], wantp) and 0 < wantp.sum() < wantp.size


def _arm_probe(obs, step=0.01, safety=0.01, max_dist=math.nan):
    ox, oy, orad = (np.ascontiguousarray(obs[:, k]) for k in range(3))
    t2 = np.array([sq_lt_threshold(r) for r in orad])
    pa = ProbeArm(obs.shape[0], p(ox), p(oy), p(orad), p(t2), step, safety,
                  sq_lt_threshold(safety), max_dist)
    pa._keep = (ox, oy, orad, t2)
    return pa


@pytest.mark.parametrize("max_dist", [math.nan, 0.6])
def test_arm_bodies_match_oracle(probe, max_dist):
    rng = np.random.default_rng(8)
    N, M = 6, 3
    rads, obs, base = make_arm_instance(N, M, seed=3)
    orc = O.Oracle(O.ARM22, rads, obs, base=base, max_dist=max_dist)
    pa = _arm_probe(obs, max_dist=max_dist)
    n = 1500
    q1 = rng.uniform(0, 2 * math.pi, (n, 2))
    q2 = rng.uniform(0, 2 * math.pi, (n, 2))
    q2[:100] = q1[:100] + rng.normal(size=(100, 2)) * 0.05     # short moves
    q2[100:120] = q1[100:120]                                   # D == 0 (NaN first step)
    q3 = rng.uniform(0, 2 * math.pi, (n, 2))
    q4 = q3 + rng.normal(size=(n, 2)) * 0.3
    ai = rng.integers(0, N, n)
    aj = (ai + 1 + rng.integers(0, N - 1, n)) % N
    n_free = n_edge = n_hit = 0
    for k in range(n):
        i, j = int(ai[k]), int(aj[k])
        bi, bj = base[i], base[j]
        w = orc.connect_point(q1[k], i)
        assert bool(probe.probe_arm_connect_point(C.byref(pa), p(q1[k]), bi[0], bi[1], rads[i])) == w
        n_free += w
        w = orc.connect_edge(q1[k], q2[k], i)
        assert bool(probe.probe_arm_connect_edge(C.byref(pa), p(q1[k]), p(q2[k]), bi[0], bi[1],
                                                 rads[i])) == w, k
        n_edge += w
        w = orc.collide_static(q1[k], q3[k], i, j)
        assert bool(probe.probe_arm_collide_static(C.byref(pa), p(q1[k]), p(q3[k]), bi[0], bi[1],
                                                   rads[i], bj[0], bj[1], rads[j])) == w
        if k < 300:
            w = orc.collide_pair(q1[k], q2[k], q3[k], q4[k], i, j)
            assert bool(probe.probe_arm_collide_pair(C.byref(pa), p(q1[k]), p(q2[k]), p(q3[k]),
                                                     p(q4[k]), bi[0], bi[1], rads[i], bj[0],
                                                     bj[1], rads[j])) == w, k
            n_hit += w
        assert bits(probe.probe_arm_dist(p(q1[k]), p(q2[k]))) == bits(orc.state_dist(q1[k], q2[k]))
        mid = np.empty(2)
        probe.probe_arm_mid_status(p(q1[k]), p(q2[k]), p(mid))
        assert np.array_equal(mid.view(np.uint64), orc.mid_status(q1[k], q2[k]).view(np.uint64))
    assert 0 < n_free < n and 0 < n_hit < 300   # both verdicts exercised


class ProbeChain(C.Structure):
    _fields_ = [("model", C.c_int), ("n_obs", C.c_int), ("obs_soa", _dp), ("t2", _dp),
                ("step", C.c_double), ("safety", C.c_double), ("safety_t2", C.c_double),
                ("max_dist", C.c_double)]


@pytest.mark.parametrize("model", [O.LINE2D, O.SNAKE2D, O.ARM33, O.CAPSULE3D])
@pytest.mark.parametrize("max_dist", [math.nan, 0.5])
def test_chain_models_match_oracle(probe, model, max_dist):
    """line2d / snake2d / arm33 / capsule3d: chain_model.cuh compiled for the host against the
    oracle's per-model restatement (oracle/orc_chain.c): connect (both forms), collide (4- and
    6-arity), dist, get_mid_status -- verdicts equal, values bit-identical."""
    from helpers import make_chain_instance, chain_states
    rng = np.random.default_rng(40 + model)
    N, M = 5, 3
    ins = make_chain_instance(model, N, M, seed=model)
    orc = O.Oracle(model, ins["rads"], ins["obs"], base=ins["base"], aux=ins["aux"], max_dist=max_dist)
    sd, wd = O._D[model], O._W[model]
    step = 0.05 if model in (O.SNAKE2D, O.CAPSULE3D) else 0.01
    obs_soa = np.ascontiguousarray(ins["obs"].T)
    geoms, t2s = [], []
    for i in range(N):
        base = ins["base"][i] if ins["base"] is not None else np.zeros(3)
        ln = ins["aux"][i] if model == O.CAPSULE3D else ins["rads"][i]
        geoms.append(np.array([*np.resize(base, 3), ln, ins["rads"][i]]))
        thr = ins["obs"][:, wd] + (ins["rads"][i] if model == O.CAPSULE3D else 0.0)
        t2s.append(np.array([sq_lt_threshold(t) for t in thr]))
    n = 400
    q1 = chain_states(rng, model, n)
    q2 = chain_states(rng, model, n, near=q1, spread=0.12)
    q2[:20] = q1[:20]                       # D == 0
    q3 = chain_states(rng, model, n, near=q1, spread=0.1)
    q4 = chain_states(rng, model, n, near=q3, spread=0.12)
    ai = rng.integers(0, N, n)
    aj = (ai + 1 + rng.integers(0, N - 1, n)) % N
    out = np.empty(8)
    tally = dict(free=0, edge=0, hit4=0, hit6=0)
    for k in range(n):
        i, j = int(ai[k]), int(aj[k])
        pc = ProbeChain(model, M, p(obs_soa), p(t2s[i]), step, 0.01, sq_lt_threshold(0.01), max_dist)
        thr = ins["rads"][i] + ins["rads"][j] if model == O.CAPSULE3D else 0.01
        T2 = sq_lt_threshold(thr)
        args = lambda what, a, b=None, c=None, d=None: probe.probe_chain(
            what, C.byref(pc), geoms[i].ctypes.data, geoms[j].ctypes.data, a.ctypes.data,
            None if b is None else b.ctypes.data, None if c is None else c.ctypes.data,
            None if d is None else d.ctypes.data, thr, T2, out.ctypes.data)
        w = orc.connect_point(q1[k], i)
        assert bool(args(0, q1[k])) == w, (k, "point")
        tally["free"] += w
        w = orc.connect_edge(q1[k], q2[k], i)
        assert bool(args(1, q1[k], q2[k])) == w, (k, "edge")
        tally["edge"] += w
        w = orc.collide_static(q1[k], q3[k], i, j)
        assert bool(args(2, q1[k], q3[k])) == w, (k, "static")
        tally["hit4"] += w
        if k < 150:
            w = orc.collide_pair(q1[k], q2[k], q3[k], q4[k], i, j)
            assert bool(args(3, q1[k], q2[k], q3[k], q4[k])) == w, (k, "pair")
            tally["hit6"] += w
        args(4, q1[k], q2[k])
        assert bits(out[0]) == bits(orc.state_dist(q1[k], q2[k]))
        args(5, q1[k], q2[k])
        assert np.array_equal(out[:sd].view(np.uint64), orc.mid_status(q1[k], q2[k]).view(np.uint64))
    assert 0 < tally["free"] < n and 0 < tally["hit4"] < n and 0 < tally["hit6"] < 150, tally
    if math.isnan(max_dist):
        assert 0 < tally["edge"] < n, tally


def _filter_cases(rng, d, n):
    """Segment pairs that stress the verdict filter: plain near pairs, contact offsets around the
    hand-over band, parallel / nearly parallel / collinear / identical / reversed segments,
    zero-length segments (stay moves), crossings at every angle, tiny segments, coordinates at and
    beyond the domain edge, non-finite values."""
    from helpers import near_threshold_pairs
    N = 12
    rads = np.concatenate([[1e-6, 2e-6, 5e-4], rng.uniform(0.002, 0.1, N - 3)])
    ai = rng.integers(0, N, n)
    aj = (ai + 1 + rng.integers(0, N - 1, n)) % N
    a, b, c, e = near_threshold_pairs(rng, n, rads, ai, aj, d)
    k = n // 8
    # plain near pairs
    a[:k] = rng.random((k, d))
    b[:k] = a[:k] + rng.normal(size=(k, d)) * 0.05
    c[:k] = a[:k] + rng.normal(size=(k, d)) * 0.05
    e[:k] = c[:k] + rng.normal(size=(k, d)) * 0.05
    # nearly parallel at angles 1e-12 .. 1e-1, offset around the threshold
    s = slice(k, 2 * k)
    thr = (rads[ai] + rads[aj])[s]
    u = rng.normal(size=(k, d)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = rng.normal(size=(k, d)); v -= (v * u).sum(1, keepdims=True) * u
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    ang = 10.0 ** rng.uniform(-12, -1, k) * rng.choice([0, 1, 1, 1], k)
    u2 = u * np.cos(ang)[:, None] + v * np.sin(ang)[:, None]
    a[s] = rng.uniform(0.2, 0.8, (k, d))
    b[s] = a[s] + u * rng.uniform(0.01, 0.1, (k, 1))
    off = v * (thr * (1 + rng.choice([0, 1e-9, -1e-9, 1e-3, -1e-3, 0.5, -0.5, -1.0], k)))[:, None]
    c[s] = a[s] + off + u * rng.uniform(-0.05, 0.05, (k, 1))
    e[s] = c[s] + u2 * rng.uniform(0.01, 0.1, (k, 1)) * rng.choice([1, -1], (k, 1))
    # zero-length segments (either, both), identical and reversed segments
    s = slice(2 * k, 3 * k)
    a[s] = rng.random((k, d)); b[s] = a[s] + rng.normal(size=(k, d)) * 0.04
    c[s] = a[s] + rng.normal(size=(k, d)) * 0.03; e[s] = c[s] + rng.normal(size=(k, d)) * 0.04
    kind = rng.integers(0, 5, k)
    b[s][kind == 0] = a[s][kind == 0]
    e[s][kind == 1] = c[s][kind == 1]
    m = kind == 2; b[s][m] = a[s][m]; e[s][m] = c[s][m]
    m = kind == 3; c[s][m] = a[s][m]; e[s][m] = b[s][m]
    m = kind == 4; c[s][m] = b[s][m]; e[s][m] = a[s][m]
    # crossings and near-crossings: second segment through a point of the first
    s = slice(3 * k, 4 * k)
    a[s] = rng.uniform(0.2, 0.8, (k, d)); b[s] = a[s] + rng.normal(size=(k, d)) * 0.05
    x = a[s] + (b[s] - a[s]) * rng.uniform(-0.1, 1.1, (k, 1))
    w = rng.normal(size=(k, d)) * 0.05
    lift = np.zeros((k, d))
    if d == 3:
        lift[:, 2] = rng.choice([0.0, 1e-12, 1e-3, 0.02, 0.05], k)
    c[s] = x - w * rng.uniform(-0.1, 1.0, (k, 1)) + lift
    e[s] = x + w * rng.uniform(-0.1, 1.0, (k, 1)) + lift
    # tiny segments (subnormal squares), domain edge, out of domain, non-finite
    s = slice(4 * k, 5 * k)
    a[s] = rng.random((k, d)); c[s] = a[s] + rng.normal(size=(k, d)) * 0.03
    b[s] = a[s] + rng.normal(size=(k, d)) * 10.0 ** rng.uniform(-320, -3, (k, 1))
    e[s] = c[s] + rng.normal(size=(k, d)) * 10.0 ** rng.uniform(-320, -3, (k, 1))
    s = slice(5 * k, 6 * k)
    sc = rng.choice([1.9, 2.0, 2.1, 10.0, 1e6], (k, 1))
    for arr in (a, b, c, e):
        arr[s] = (arr[s] - 0.5) * 2 * sc / np.maximum(1, np.abs(arr[s] - 0.5).max(1, keepdims=True) * 2) * \
            rng.uniform(0.9, 1.0, (k, 1))
    sp = rng.integers(5 * k, 6 * k, 64)
    a[sp[:16], 0] = np.nan; b[sp[16:32], 1] = np.inf; c[sp[32:48], 0] = -np.inf; e[sp[48:], 1] = np.nan
    return rads, ai, aj, [np.ascontiguousarray(t) for t in (a, b, c, e)]


@pytest.mark.parametrize("d", [2, 3])
def test_verdict_filter_never_contradicts_reference_arithmetic(probe, d):
    """geom.cuh: segseg_filter (the conflict-table kernel's cheap exact-verdict filter) answers
    1 / 0 only where segseg_lt -- the reference's own sequence of roundings -- says the same;
    everything else must come back undecided.  Also: it decides nearly all ordinary pairs."""
    rng = np.random.default_rng(100 + d)
    n = 400_000
    rads, ai, aj, (a, b, c, e) = _filter_cases(rng, d, n)
    thr = rads[ai] + rads[aj]
    T2 = np.array([sq_lt_threshold(t) for t in np.unique(thr)])
    T2 = T2[np.searchsorted(np.unique(thr), thr)]
    out = np.empty(n, np.int8)
    exact = np.empty(n, np.uint8)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    probe.probe_segseg_filter_batch(n, d, vp(a), vp(b), vp(c), vp(e), vp(thr), vp(T2), vp(out), vp(exact))
    decided = out >= 0
    bad = decided & (out != exact.astype(np.int8))
    assert not bad.any(), (int(bad.sum()), np.flatnonzero(bad)[:10])
    # the exact path agrees with the oracle on a sample (pins `exact` itself)
    for k in rng.integers(0, n, 3000):
        assert bool(exact[k]) == (O.dist_ss(a[k], b[k], c[k], e[k]) < thr[k]), k
    k8 = n // 8
    plain = decided[:k8].mean()
    # (1.5 % of the pairs combine the two sub-2^-18 radii, for which the filter is switched off)
    print('decided: plain %.4f, all %.4f' % (plain, decided.mean()))
    assert plain > (0.98 if d == 2 else 0.5), plain
    # both verdicts and the undecided answer all occur
    assert (out == 1).any() and (out == 0).any() and (out == -1).any()


@pytest.mark.parametrize("d", [2, 3])
def test_fp32_broad_phase_is_one_sided(probe, d):
    """geom.cuh: bound32_far (single-precision capsule bound of the conflict-table kernel) may
    only reject pairs whose true distance exceeds r_i + r_j by a margin: checked against a
    float128 evaluation of the true segment distance on pairs placed around the rejection
    boundary, at every segment length the domain allows, plus the filter's adversarial set."""
    rng = np.random.default_rng(200 + d)
    n = 300_000
    ri, rj = rng.uniform(1e-4, 0.05, n), rng.uniform(1e-4, 0.05, n)
    # segment 1 anywhere in [-2, 2]^d, random length; segment 2 placed so that the true distance
    # is (r_i + r_j) * (1 + tiny .. 1) -- i.e. just beyond contact, where a sloppy bound errs
    L1 = 10.0 ** rng.uniform(-4, 0.3, n) * rng.choice([0, 1, 1, 1], n)
    L2 = 10.0 ** rng.uniform(-4, 0.3, n) * rng.choice([0, 1, 1, 1], n)
    u = rng.normal(size=(n, d)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    v = rng.normal(size=(n, d)); v -= (v * u).sum(1, keepdims=True) * u
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    mid = rng.uniform(-1.0, 1.0, (n, d))
    a = mid - u * (L1 / 2)[:, None]
    b = mid + u * (L1 / 2)[:, None]
    gap = (ri + rj) * (1 + 10.0 ** rng.uniform(-7, 0, n))
    along = rng.uniform(-0.6, 0.6, n) * L1        # closest point somewhere along segment 1
    foot = mid + u * along[:, None] + v * gap[:, None]
    w = rng.normal(size=(n, d)); w -= (w * v).sum(1, keepdims=True) * v
    nw = np.linalg.norm(w, axis=1, keepdims=True); w /= np.where(nw > 0, nw, 1)
    s0 = rng.uniform(0, 1, n)
    c = foot - w * (L2 * s0)[:, None]
    e = foot + w * (L2 * (1 - s0))[:, None]
    keep = (np.abs(a).max(1) <= 2) & (np.abs(b).max(1) <= 2) & (np.abs(c).max(1) <= 2) & (np.abs(e).max(1) <= 2)
    a, b, c, e = (np.ascontiguousarray(x[keep]) for x in (a, b, c, e))
    ri, rj = np.ascontiguousarray(ri[keep]), np.ascontiguousarray(rj[keep])
    m = a.shape[0]
    out = np.empty(m, np.uint8)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    probe.probe_bound32_batch(m, d, vp(a), vp(b), vp(c), vp(e), vp(ri), vp(rj), vp(out))
    assert out.any() and not out.all()
    # every rejected pair must be a "no collision" of the reference arithmetic ...
    rej = np.flatnonzero(out)
    for k in rej[:: max(1, rej.size // 20000)]:
        assert not (O.dist_ss(a[k], b[k], c[k], e[k]) < ri[k] + rj[k]), k
    # ... with the promised margin (true distance by dense sampling in long double: an upper
    # bound of the true distance, so `>=` here is implied by the claim)
    ld = np.longdouble
    ts = np.linspace(0, 1, 65).astype(ld)
    for k in rej[:: max(1, rej.size // 3000)]:
        P = a[k].astype(ld)[None, :] + ts[:, None] * (b[k] - a[k]).astype(ld)[None, :]
        Q = c[k].astype(ld)[None, :] + ts[:, None] * (e[k] - c[k]).astype(ld)[None, :]
        dmin = np.sqrt(((P[:, None, :] - Q[None, :, :]) ** 2).sum(-1)).min()
        assert dmin > ri[k] + rj[k] + 9.0e-7, (k, float(dmin), ri[k] + rj[k])
    # the filter's adversarial set: out-of-domain and non-finite segments are never rejected
    rads, ai, aj, (a2, b2, c2, e2) = _filter_cases(rng, d, 80_000)
    out2 = np.empty(80_000, np.uint8)
    r_i, r_j = np.ascontiguousarray(rads[ai]), np.ascontiguousarray(rads[aj])
    probe.probe_bound32_batch(80_000, d, vp(a2), vp(b2), vp(c2), vp(e2), vp(r_i), vp(r_j), vp(out2))
    bad = [k for k in np.flatnonzero(out2) if O.dist_ss(a2[k], b2[k], c2[k], e2[k]) < r_i[k] + r_j[k]]
    assert not bad, bad[:10]
    fin = np.isfinite(a2).all(1) & np.isfinite(b2).all(1) & np.isfinite(c2).all(1) & np.isfinite(e2).all(1)
    big = (np.abs(a2).max(1) > 2) | (np.abs(b2).max(1) > 2) | (np.abs(c2).max(1) > 2) | (np.abs(e2).max(1) > 2)
    assert not out2[~fin | big].any()
