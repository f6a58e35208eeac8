"""tests/golden/ref_pin_point2d_n6.json is what julia/ref_bench.jl runs the REAL reference on (on a
machine that has Julia) to pin the oracle.  Here: the committed fixture is exactly what the
generator writes today (the C oracle has not drifted), and the literal pure-Python transliteration
of the reference (oracle/pyref.py) -- independent of the C code -- gives the same neighbour lists,
collide verdicts and connect verdicts."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "ref_pin_point2d_n6.json")


def _load():
    with open(FIX) as f:
        return json.load(f)


def test_fixture_is_what_the_generator_writes():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import export_ref_fixture as E
    fx = _load()
    again = E.make(6, 60, 0.25, 8, 7, 300, 0)
    again = json.loads(json.dumps(again))      # same float -> text -> float path as the file
    for k in ("rads", "obstacles", "vertices", "neighbors", "table", "points"):
        assert again[k] == fx[k], k


def test_python_transliteration_agrees_with_the_fixture():
    from oracle import pyref as R
    fx = _load()
    inst = R.PointInstance(fx["rads"], fx["obstacles"])
    V = fx["vertices"]
    for i in range(fx["n_agents"]):
        assert inst.prm_neighbors(V[i], fx["rad"], i) == fx["neighbors"][i], "agent %d" % i
    tb = fx["table"]
    n_hit = 0
    for r, (i, vf, vt) in enumerate(zip(tb["row_agent"], tb["row_vfrom"], tb["row_vto"])):
        got = []
        for j, cf, ct in zip(tb["col_agent"], tb["col_vfrom"], tb["col_vto"]):
            got.append("1" if j != i and inst.collide_pair(V[i][vf], V[i][vt], V[j][cf], V[j][ct], i, j) else "0")
        assert "".join(got) == tb["verdict"][r], "row %d" % r
        n_hit += got.count("1")
    assert n_hit > 0
    pt = fx["points"]
    got = [int(inst.connect_point(q, a)) for a, q in zip(pt["agent"], pt["q"])]
    assert got == pt["connect"] and 0 < sum(got) < len(got)


def test_julia_script_uses_only_the_reference_api():
    src = open(os.path.join(ROOT, "julia", "ref_bench.jl")).read()
    for name in ("gen_connect", "gen_collide", "MRMP.Solvers.PRM!", "oracle_pinned"):
        assert name in src
    assert "ccall" not in src      # the reference arm must not touch this repo's library
