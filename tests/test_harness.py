"""The eval harness speaks the reference's formats: YAML experiment configs with k=v overrides
(scripts/eval.jl:177-213) and the result.csv column schema (eval.jl:67-83)."""
import csv
import glob
import os

import pytest

from sssp_b200 import harness as H

CFG = """
root: ../data/exp/point2d_tiny
seed_offset: 3
time_limit: 30
num_instances: 2
instance:
  _target_: MRMP.gen_random_instance_StatePoint2D
  N: 3
  rad_min: 0.03
  rad_max: 0.05
  num_obs: 4
  rad_obs_min: 0.05
  rad_obs_max: 0.08
solvers:
  - _target_: MRMP.SSSP
    epsilon: 0.3
    init_min_dist_thread: 0.05
  - _target_: MRMP.PP
    num_vertices: 100
"""


def test_config_overrides_follow_eval_jl(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text(CFG)
    c = H.load_config(str(p), ["instance.N=5", "time_limit=2.5", "root=out", "solvers=x"])
    assert c["instance"]["N"] == 5 and isinstance(c["instance"]["N"], int)
    assert c["time_limit"] == 2.5 and c["root"] == "out" and c["solvers"] == "x"
    with pytest.raises(KeyError):
        H.load_config(str(p), ["nosuch.key=1"])
    assert H.parse_value("12") == 12 and H.parse_value("1e-3") == 1e-3 and H.parse_value("~") == "~"
    # the reference's own experiment configs parse (they are read from the survey's copy of the
    # schema, not from /root/reference, which does not exist on the GPU box)
    assert H.RESULT_COLUMNS[:6] == ["instance", "N", "num_obs", "solver", "solver_index", "solved"]


@pytest.mark.gpu
def test_harness_runs_a_config_and_writes_result_csv(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text(CFG)
    config = H.load_config(str(p), [])
    rows = H.run(config, out_dir=str(tmp_path / "exp"))
    assert len(rows) == 4                                    # 2 instances x 2 solver entries
    assert [r["solver"] for r in rows] == ["MRMP.SSSP", "MRMP.PP"] * 2
    assert all(r["solved"] for r in rows if r["solver"] == "MRMP.SSSP")
    assert all(not r["solved"] and r["elapsed_planning"] == 0 for r in rows if r["solver"] == "MRMP.PP")
    assert all(r["makespan_original"] > 0 and r["sum_of_cost_original"] >= r["makespan_original"]
               for r in rows if r["solved"])
    out = glob.glob(str(tmp_path / "exp" / "*" / "result.csv"))
    assert len(out) == 1 and os.path.exists(os.path.join(os.path.dirname(out[0]), "config.yaml"))
    with open(out[0]) as f:
        table = list(csv.DictReader(f))
    assert list(table[0].keys()) == H.RESULT_COLUMNS and table[0]["solved"] == "true"
    # deterministic: same config, same rows (apart from wall-clock columns)
    rows2 = H.run(config)
    strip = lambda r: {k: v for k, v in r.items() if not k.startswith("elapsed")}
    assert [strip(r) for r in rows] == [strip(r) for r in rows2]
