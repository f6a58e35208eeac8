"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  See oracle/mrmp_oracle.h for scope and pinning.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

POINT2D, POINT3D, ARM22, LINE2D, SNAKE2D, ARM33, CAPSULE3D = 0, 1, 2, 3, 4, 5, 6
_D = {POINT2D: 2, POINT3D: 3, ARM22: 2, LINE2D: 3, SNAKE2D: 6, ARM33: 6, CAPSULE3D: 6}   # state dim
_W = {POINT2D: 2, POINT3D: 3, ARM22: 2, LINE2D: 2, SNAKE2D: 2, ARM33: 3, CAPSULE3D: 3}   # workspace dim

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile oracle/*.c with the committed Makefile (gcc only, no reference build)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.orc_ctx_create.restype = C.c_void_p
    lib.orc_ctx_create.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double,
                                   C.c_double, C.c_double]
    lib.orc_ctx_create2.restype = C.c_void_p
    lib.orc_ctx_create2.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp, C.c_double,
                                    C.c_double, C.c_double]
    lib.orc_ctx_destroy.argtypes = [C.c_void_p]
    for f in ("orc_norm", "orc_dist_pp", "orc_dist_sp", "orc_dist_ss", "orc_diff_angles",
              "orc_sin", "orc_cos", "orc_atan2", "orc_state_dist", "orc_dot"):
        getattr(lib, f).restype = C.c_double
    lib.orc_norm.argtypes = [_dp, C.c_int]
    lib.orc_dot.argtypes = [_dp, _dp, C.c_int]
    lib.orc_dist_pp.argtypes = [_dp, _dp, C.c_int]
    lib.orc_dist_sp.argtypes = [_dp, _dp, _dp, C.c_int]
    lib.orc_dist_ss.argtypes = [_dp, _dp, _dp, _dp, C.c_int]
    lib.orc_diff_angles.argtypes = [C.c_double, C.c_double]
    lib.orc_sin.argtypes = [C.c_double]
    lib.orc_cos.argtypes = [C.c_double]
    lib.orc_atan2.argtypes = [C.c_double, C.c_double]
    lib.orc_state_dist.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_mid_status.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib.orc_step_schedule.restype = C.c_int
    lib.orc_step_schedule.argtypes = [C.c_double, C.c_double, _dp, C.c_int]
    lib.orc_connect_point.argtypes = [C.c_void_p, _dp, C.c_int]
    lib.orc_connect_edge.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    lib.orc_collide_pair.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, C.c_int]
    lib.orc_collide_static.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int]
    lib.orc_collide_configs.argtypes = [C.c_void_p, _dp, _dp, _lp]
    lib.orc_collide_config_move.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    lib.orc_connect_points_batch.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _bp, C.c_int]
    lib.orc_connect_edges_batch.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _dp, _bp, C.c_int]
    lib.orc_collide_pairs_batch.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp, _dp, _dp, _dp,
                                            _bp, C.c_int]
    lib.orc_collide_static_pairs_batch.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp, _dp,
                                                   _bp, C.c_int]
    lib.orc_collide_configs_batch.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _bp, _lp, C.c_int]
    lib.orc_collide_config_moves_batch.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int64, _dp,
                                                   _bp, C.c_int]
    lib.orc_collide_edges_indexed_batch.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _ip, _ip,
                                                    _ip, _ip, _dp, C.c_int64, _bp, C.c_int]
    lib.orc_collide_cross.argtypes = [C.c_void_p, C.c_int64, _ip, _dp, _dp, C.c_int64, _ip, _dp, _dp,
                                      C.c_int, C.c_void_p, C.c_int]
    lib.orc_collide_cross_reduce.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_int]
    lib.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    lib.orc_sample_state.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, _dp]
    lib.orc_sample_free.restype = C.c_int64
    lib.orc_sample_free.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, _dp, C.c_int64]
    lib.orc_prm_build_agent.restype = C.c_int64
    lib.orc_prm_build_agent.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp, _ip, _ip,
                                        C.c_int64]
    lib.orc_prm_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip, _ip,
                                  C.c_int64, _lp, C.c_int]
    lib.orc_steering.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, C.c_double, _dp]
    lib.orc_conn.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double]
    lib.orc_distance_table.argtypes = [C.c_void_p, C.c_int, _dp, _ip, _ip, C.c_int, _dp]
    lib.orc_sssp_expand.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp,
                                    C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                    C.c_void_p, _dp, C.c_int, _ip, C.c_int, _ip, _bp,
                                    C.POINTER(C.c_int)]
    return lib


_LIBS: dict = {}


def lib(dot_fma: bool = False) -> C.CDLL:
    name = "libmrmp_oracle_dotfma.so" if dot_fma else "libmrmp_oracle.so"
    if name not in _LIBS:
        _LIBS[name] = _load(name)
    return _LIBS[name]


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def host_cores() -> int:
    """Cores this process may run on.  The batch drivers take an explicit thread count
    (`num_threads` clause), so this is what the reference arm passes: OMP_NUM_THREADS -- which
    torchrun sets to 1 in every worker -- must not decide how many host cores the CPU arm uses."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------- primitives
def dist_pp(a, b) -> float:
    a, b = _f64(a), _f64(b)
    return lib().orc_dist_pp(_p(a, _dp), _p(b, _dp), a.size)


def dist_sp(a, b, c, dot_fma=False) -> float:
    a, b, c = _f64(a), _f64(b), _f64(c)
    return lib(dot_fma).orc_dist_sp(_p(a, _dp), _p(b, _dp), _p(c, _dp), a.size)


def dist_ss(p11, p12, p21, p22, dot_fma=False) -> float:
    p11, p12, p21, p22 = _f64(p11), _f64(p12), _f64(p21), _f64(p22)
    return lib(dot_fma).orc_dist_ss(_p(p11, _dp), _p(p12, _dp), _p(p21, _dp), _p(p22, _dp),
                                    p11.size)


def norm(v) -> float:
    v = _f64(v)
    return lib().orc_norm(_p(v, _dp), v.size)


def diff_angles(t1: float, t2: float) -> float:
    return lib().orc_diff_angles(t1, t2)


def step_schedule(step: float, D: float) -> np.ndarray:
    out = np.empty(8192)
    n = lib().orc_step_schedule(step, D, _p(out, _dp), out.size)
    assert n <= out.size
    return out[:n].copy()


def philox4x32(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32(*[int(x) for x in ctr], int(key[0]), int(key[1]), out)
    return [int(x) for x in out]


# ------------------------------------------------------------------- context
class Oracle:
    """CPU restatement of gen_connect / gen_collide closures for one instance."""

    def __init__(self, model, rads, obstacles, base=None, step_dist=None, safety_dist=0.01,
                 max_dist=math.nan, dot_fma=False, aux=None):
        self.model = int(model)
        self.d = _D[self.model]
        self.wd = _W[self.model]
        self.rads = _f64(rads).ravel()
        self.N = self.rads.size
        self.obs = _f64(obstacles).reshape(-1, self.wd + 1) if np.size(obstacles) else \
            np.zeros((0, self.wd + 1))
        self.base = None if base is None else _f64(base).reshape(self.N, self.wd)
        self.aux = None if aux is None else _f64(aux).ravel()
        self._lib = lib(dot_fma)
        if max_dist is None:
            max_dist = math.nan
        if step_dist is None:   # STEP_DIST (MRMP.jl:65); snake2d.jl:3 / capsule3d.jl:3 use 0.05
            step_dist = 0.05 if self.model in (SNAKE2D, CAPSULE3D) else 0.01
        self._h = self._lib.orc_ctx_create2(
            self.model, self.N, _p(self.rads, _dp),
            _p(self.base, _dp) if self.base is not None else None,
            _p(self.aux, _dp) if self.aux is not None else None,
            self.obs.shape[0], _p(self.obs, _dp) if self.obs.size else None,
            step_dist, safety_dist, max_dist)
        if not self._h:
            raise ValueError("orc_ctx_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.orc_ctx_destroy(h)

    # singles
    def state_dist(self, q1, q2) -> float:
        q1, q2 = _f64(q1), _f64(q2)
        return self._lib.orc_state_dist(self._h, _p(q1, _dp), _p(q2, _dp))

    def mid_status(self, p, q) -> np.ndarray:
        p, q = _f64(p), _f64(q)
        out = np.empty(self.d)
        self._lib.orc_mid_status(self._h, _p(p, _dp), _p(q, _dp), _p(out, _dp))
        return out

    def connect_point(self, q, i) -> bool:
        q = _f64(q)
        return bool(self._lib.orc_connect_point(self._h, _p(q, _dp), int(i)))

    def connect_edge(self, qf, qt, i) -> bool:
        qf, qt = _f64(qf), _f64(qt)
        return bool(self._lib.orc_connect_edge(self._h, _p(qf, _dp), _p(qt, _dp), int(i)))

    def collide_pair(self, qif, qit, qjf, qjt, i, j) -> bool:
        a, b, c, d = _f64(qif), _f64(qit), _f64(qjf), _f64(qjt)
        return bool(self._lib.orc_collide_pair(self._h, _p(a, _dp), _p(b, _dp), _p(c, _dp),
                                               _p(d, _dp), int(i), int(j)))

    def collide_static(self, qi, qj, i, j) -> bool:
        a, b = _f64(qi), _f64(qj)
        return bool(self._lib.orc_collide_static(self._h, _p(a, _dp), _p(b, _dp), int(i), int(j)))

    def collide_configs(self, Qf, Qt):
        a, b = _f64(Qf), _f64(Qt)
        fp = C.c_int64(-1)
        r = self._lib.orc_collide_configs(self._h, _p(a, _dp), _p(b, _dp), C.byref(fp))
        return bool(r), int(fp.value)

    def collide_config_move(self, Q, q_to, i) -> bool:
        a, b = _f64(Q), _f64(q_to)
        return bool(self._lib.orc_collide_config_move(self._h, _p(a, _dp), _p(b, _dp), int(i)))

    # batches
    def connect_points(self, agent, q, nthreads=1) -> np.ndarray:
        agent, q = _i32(agent), _f64(q)
        out = np.empty(agent.size, np.uint8)
        self._lib.orc_connect_points_batch(self._h, agent.size, _p(agent, _ip), _p(q, _dp),
                                           _p(out, _bp), nthreads)
        return out

    def connect_edges(self, agent, qf, qt, nthreads=1) -> np.ndarray:
        agent, qf, qt = _i32(agent), _f64(qf), _f64(qt)
        out = np.empty(agent.size, np.uint8)
        self._lib.orc_connect_edges_batch(self._h, agent.size, _p(agent, _ip), _p(qf, _dp),
                                          _p(qt, _dp), _p(out, _bp), nthreads)
        return out

    def collide_pairs(self, ai, aj, qif, qit, qjf, qjt, nthreads=1) -> np.ndarray:
        ai, aj = _i32(ai), _i32(aj)
        a, b, c, d = _f64(qif), _f64(qit), _f64(qjf), _f64(qjt)
        out = np.empty(ai.size, np.uint8)
        self._lib.orc_collide_pairs_batch(self._h, ai.size, _p(ai, _ip), _p(aj, _ip), _p(a, _dp),
                                          _p(b, _dp), _p(c, _dp), _p(d, _dp), _p(out, _bp),
                                          nthreads)
        return out

    def collide_static_pairs(self, ai, aj, qi, qj, nthreads=1) -> np.ndarray:
        ai, aj, a, b = _i32(ai), _i32(aj), _f64(qi), _f64(qj)
        out = np.empty(ai.size, np.uint8)
        self._lib.orc_collide_static_pairs_batch(self._h, ai.size, _p(ai, _ip), _p(aj, _ip),
                                                 _p(a, _dp), _p(b, _dp), _p(out, _bp), nthreads)
        return out

    def collide_configs_batch(self, Qf, Qt, nthreads=1):
        a, b = _f64(Qf), _f64(Qt)
        n = a.size // (self.N * self.d)
        out = np.empty(n, np.uint8)
        fp = np.empty(n, np.int64)
        self._lib.orc_collide_configs_batch(self._h, n, _p(a, _dp), _p(b, _dp), _p(out, _bp),
                                            _p(fp, _lp), nthreads)
        return out, fp

    def collide_config_moves(self, agent_i, Q, q_to, nthreads=1) -> np.ndarray:
        Q, q_to = _f64(Q), _f64(q_to)
        n = q_to.size // self.d
        out = np.empty(n, np.uint8)
        self._lib.orc_collide_config_moves_batch(self._h, int(agent_i), _p(Q, _dp), n,
                                                 _p(q_to, _dp), _p(out, _bp), nthreads)
        return out

    def collide_edges_indexed(self, ai, aj, vif, vit, vjf, vjt, nodes, nv_stride, nthreads=1):
        ai, aj, vif, vit, vjf, vjt = map(_i32, (ai, aj, vif, vit, vjf, vjt))
        nodes = _f64(nodes)
        out = np.empty(ai.size, np.uint8)
        self._lib.orc_collide_edges_indexed_batch(
            self._h, ai.size, _p(ai, _ip), _p(aj, _ip), _p(vif, _ip), _p(vit, _ip), _p(vjf, _ip),
            _p(vjt, _ip), _p(nodes, _dp), int(nv_stride), _p(out, _bp), nthreads)
        return out

    def collide_cross(self, ra, rf, rt, ca, cf, ct, group_size=0, nthreads=1) -> np.ndarray:
        ra, ca = _i32(ra), _i32(ca)
        rf, rt, cf, ct = _f64(rf), _f64(rt), _f64(cf), _f64(ct)
        bits = -(-ca.size // group_size) if group_size > 0 else ca.size
        out = np.zeros((ra.size, (bits + 31) // 32), np.uint32)
        self._lib.orc_collide_cross(self._h, ra.size, _p(ra, _ip), _p(rf, _dp), _p(rt, _dp),
                                    ca.size, _p(ca, _ip), _p(cf, _dp), _p(ct, _dp),
                                    int(group_size), out.ctypes.data_as(C.c_void_p), nthreads)
        return out

    def collide_cross_reduce(self, ra, rf, rt, ca, cf, ct, mode, group_size=0, col_group=None,
                             n_groups=None, row_key=None, col_key=None, cbs=False, nthreads=1):
        """mode "count" (cbs.jl:336-347) / "first" (smoother.jl:78-106) -> int32 [n_rows][n_groups]."""
        ra, ca = _i32(ra), _i32(ca)
        rf, rt, cf, ct = _f64(rf), _f64(rt), _f64(cf), _f64(ct)
        cg = _i32(col_group) if col_group is not None else None
        rk = _i32(row_key) if row_key is not None else None
        ck = _i32(col_key) if col_key is not None else None
        if n_groups is None:
            n_groups = int(cg.max()) + 1 if cg is not None else -(-ca.size // group_size)
        out = np.zeros((ra.size, n_groups), np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self._lib.orc_collide_cross_reduce(self._h, ra.size, vp(ra), vp(rf), vp(rt), vp(rk), ca.size,
                                           vp(ca), vp(cf), vp(ct), vp(cg), vp(ck), int(group_size),
                                           int(n_groups), {"count": 2, "first": 3}[mode], int(bool(cbs)),
                                           vp(out), nthreads)
        return out

    # sampler / roadmap
    def sample_state(self, seed, agent, t) -> np.ndarray:
        q = np.empty(self.d)
        self._lib.orc_sample_state(self._h, int(seed), int(agent), int(t), _p(q, _dp))
        return q

    def sample_free(self, agent, seed, n_wanted, max_tries=None):
        """First n_wanted free samples in try order -> (q, tries).  Gives up (RuntimeError)
        after 16384 tries per wanted sample, like the library (an agent without any free
        posture would otherwise loop forever)."""
        q = np.empty((n_wanted, self.d))
        explicit = max_tries is not None
        max_tries = int(max_tries) if explicit else (1 << 14) * max(int(n_wanted), 1)
        tried = self._lib.orc_sample_free(self._h, int(agent), int(seed), int(n_wanted),
                                          _p(q, _dp), max_tries)
        if tried < 0 and not explicit:
            raise RuntimeError("oracle: free-space sampling gave up for agent %d" % agent)
        return q, int(tried)

    def prm_build(self, agent0, nodes, rad, cap=None, nthreads=1):
        """nodes [n][nv][d]; rad scalar/array (nan == nothing) -> row_ptr [n][nv+1], col lists"""
        nodes = _f64(nodes)
        n, nv = nodes.shape[0], nodes.shape[1]
        rad = np.broadcast_to(_f64(np.nan if rad is None else rad), (n,)).copy()
        cap = int(cap if cap is not None else nv * (nv - 1))
        row_ptr = np.empty((n, nv + 1), np.int32)
        col = np.empty((n, max(cap, 1)), np.int32)
        nnz = np.empty(n, np.int64)
        self._lib.orc_prm_build(self._h, int(agent0), n, nv, _p(rad, _dp), _p(nodes, _dp),
                                _p(row_ptr, _ip), _p(col, _ip), cap, _p(nnz, _lp), nthreads)
        return row_ptr, col, nnz

    def conn(self, agent, qf, qt, epsilon=math.nan) -> bool:
        qf, qt = _f64(qf), _f64(qt)
        eps = math.nan if epsilon is None else epsilon
        return bool(self._lib.orc_conn(self._h, int(agent), _p(qf, _dp), _p(qt, _dp), eps))

    def steering(self, agent, q_rand, q_near, depth, epsilon=math.nan) -> np.ndarray:
        q_rand, q_near = _f64(q_rand), _f64(q_near)
        eps = math.nan if epsilon is None else epsilon
        out = np.empty(self.d)
        self._lib.orc_steering(self._h, int(agent), _p(q_rand, _dp), _p(q_near, _dp), int(depth),
                               eps, _p(out, _dp))
        return out

    def sssp_expand(self, agent, nodes, v_from, q_rand, depth, epsilon, thr, Q, old_nbrs,
                    slow=False):
        """One SSSP node expansion in C (expand! + successor filter).  nodes [nv][d] is not
        modified; returns (nodes_after, out_bits, in_bits, succ_ids, succ_collide)."""
        nodes, q_rand, Q = _f64(nodes), _f64(q_rand).reshape(-1, self.d), _f64(Q)
        nv, K = nodes.shape[0], q_rand.shape[0]
        buf = np.zeros((nv + K, self.d))
        buf[:nv] = nodes
        words = (nv + K + 31) // 32
        ob, ib = np.zeros((K, words), np.uint32), np.zeros((K, words), np.uint32)
        old = _i32(old_nbrs)
        succ, hit = np.empty(old.size + K + 1, np.int32), np.empty(old.size + K + 1, np.uint8)
        ns = C.c_int(0)
        eps = math.nan if epsilon is None else epsilon
        n_new = self._lib.orc_sssp_expand(
            self._h, int(agent), _p(buf, _dp), nv, int(v_from), K, _p(q_rand, _dp), int(depth), eps,
            float(thr), words, ob.ctypes.data_as(C.c_void_p), ib.ctypes.data_as(C.c_void_p),
            _p(Q, _dp), old.size, _p(old, _ip) if old.size else None, int(slow), _p(succ, _ip),
            _p(hit, _bp), C.byref(ns))
        return buf[: nv + n_new], ob[:n_new], ib[:n_new], succ[: ns.value], hit[: ns.value]

    def distance_table(self, nodes, row_ptr, col_idx, goal=1) -> np.ndarray:
        nodes, row_ptr, col_idx = _f64(nodes), _i32(row_ptr), _i32(col_idx)
        nv = row_ptr.size - 1
        out = np.empty(nv)
        self._lib.orc_distance_table(self._h, nv, _p(nodes, _dp), _p(row_ptr, _ip),
                                     _p(col_idx, _ip), int(goal), _p(out, _dp))
        return out
