"""Experiment harness in the reference's formats (scripts/eval.jl + scripts/config/eval/*.yaml).

Lets the reference's experiment configs drive the GPU path unchanged: the same YAML schema
(`instance._target_ = MRMP.gen_random_instance_State*`, `solvers[*]._target_`, `time_limit`,
`seed_offset`, `num_instances`), the same `key.sub=value` command-line overrides
(eval.jl:186-213) and the same `result.csv` columns (eval.jl:67-83).

What differs, by necessity: instances are drawn with numpy's PCG64 instead of Julia's RNG
(the procedure is gen_random_instance, utils.jl:204-245, without its wall-clock retry); only
solvers built on this path are run (`MRMP.SSSP`) -- other `_target_`s produce a row with
`solved = false` and `elapsed_planning = 0` so that the table keeps the reference's shape.
`smoothing` (src/smoother.jl, eval.jl:64) runs on the batched path (sssp_b200/smoother.py) and
fills the `*_refined` columns and `elapsed_refinement`.

`planning(req)` answers one request of the reference's planning server (scripts/server.jl:84-122)
in its JSON format -- {"status": "success", "instructions": [...]} -- without the WebSocket
transport, which is not part of the path.
"""
from __future__ import annotations

import csv
import datetime
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import closures as CL
from . import lib as L
from .solvers import SSSP
from .smoother import get_instructions, smoothing
from .validation import is_valid_instance, validate

RESULT_COLUMNS = ["instance", "N", "num_obs", "solver", "solver_index", "solved", "elapsed_planning",
                  "elapsed_refinement", "elapsed_total", "sum_of_cost_original", "makespan_original",
                  "sum_of_cost_refined", "makespan_refined"]          # eval.jl:67-83

_STATE_BY_NAME = {t.__name__: t for t in CL._MODEL_OF}
_NPOS = {L.MODEL_POINT2D: 2, L.MODEL_POINT3D: 3, L.MODEL_ARM22: 0, L.MODEL_LINE2D: 2,
         L.MODEL_SNAKE2D: 2, L.MODEL_ARM33: 0, L.MODEL_CAPSULE3D: 3}


def parse_value(val: str):
    """eval.jl:203-210: Int64, then Float64, else the string."""
    for cast in (int, float):
        try:
            return cast(val)
        except ValueError:
            pass
    return val


def load_config(path: str, overrides: Sequence[str] = ()) -> dict:
    """YAML file + `a.b.c=value` overrides (eval.jl:177-213)."""
    import yaml
    with open(path) as f:
        config = yaml.safe_load(f)
    for arg in overrides:
        keys, val = arg.split("=", 1)
        *parents, last = keys.split(".")
        node = config
        for k in parents:
            if k not in node:
                raise KeyError(f"{path} does not have key {'.'.join(parents)}")
            node = node[k]
        node[last] = parse_value(val)
    return config


def gen_random_instance(State, rng: np.random.Generator, *, N_min=2, N_max=8, N=None, num_obs_min=0,
                        num_obs_max=10, num_obs=None, rad=0.025, rad_obs=0.05, rad_min=None,
                        rad_max=None, rad_obs_min=None, rad_obs_max=None, max_tries=20000,
                        device: int = 0, **_ignored):
    """gen_random_instance (utils.jl:204-245) + gen_config_init_goal (:160-201): obstacles,
    radii, then per agent a start and a goal that are free and collide with no start / goal
    chosen before.  Validity comes from the GPU closures.  Returns
    (config_init, config_goal, obstacles, ins_params)."""
    model = CL._MODEL_OF[State]
    d, wd = L.STATE_DIM[model], L.WORK_DIM[model]
    N = int(rng.integers(N_min, N_max + 1)) if N is None else int(N)
    num_obs = int(rng.integers(num_obs_min, num_obs_max + 1)) if num_obs is None else int(num_obs)
    rad_min, rad_max = (rad if rad_min is None else rad_min), (rad if rad_max is None else rad_max)
    ro_min = rad_obs if rad_obs_min is None else rad_obs_min
    ro_max = rad_obs if rad_obs_max is None else rad_obs_max
    while True:
        obstacles = np.array([list(rng.random(wd)) + [rng.random() * (ro_max - ro_min) + ro_min]
                              for _ in range(num_obs)]).reshape(-1, wd + 1)
        rads = np.array([rng.random() * (rad_max - rad_min) + rad_min for _ in range(N)])
        if model in (L.MODEL_ARM22, L.MODEL_ARM33):     # gen_random_instance_StateArm22, arm22.jl:296
            ins_params = ([list(rng.random(wd)) for _ in range(N)], rads)
        elif model == L.MODEL_CAPSULE3D:
            ins_params = (rads, np.array([rng.random() * 0.1 + 0.05 for _ in range(N)]))
        else:
            ins_params = (rads,)
        tag = State(*([0.0] * d))
        connect = CL.gen_connect(tag, obstacles, *ins_params, device=device)
        collide = CL.gen_collide(tag, *ins_params, device=device)
        init: List[np.ndarray] = []
        goal: List[np.ndarray] = []
        ok = True
        for i in range(N):
            for chosen in (init, goal):
                for _ in range(max_tries):
                    q = rng.random(d)
                    q[_NPOS[model]:] *= 2 * np.pi
                    if not connect.points([i], [q])[0]:
                        continue
                    if chosen and collide.static_pairs([i] * len(chosen), list(range(len(chosen))),
                                                       [q] * len(chosen), chosen).any():
                        continue
                    chosen.append(q)
                    break
                else:
                    ok = False
                    break
            if not ok:
                break
        if ok:
            return ([State(*q) for q in init], [State(*q) for q in goal], obstacles, ins_params)


def get_solution_cost(ctx: CL.Context, solution) -> Optional[Dict[str, float]]:
    """smoother.jl:304-330: sum of costs and makespan of a synchronous solution."""
    if solution is None:
        return None
    N, T = len(solution[0]), len(solution)
    last = [T] * N
    for i in range(N):
        for t in range(T - 1, 0, -1):           # Julia t = T-1 .. 1 (1-based)
            if tuple(solution[t - 1][i].q) != tuple(solution[t][i].q):
                break
            last[i] = t
    makespan = sum_of_cost = 0.0
    for t in range(2, T + 1):
        c = float(np.max(ctx.state_dist([v.q for v in solution[t - 2]], [v.q for v in solution[t - 1]])))
        sum_of_cost += c * sum(1 for e in last if t <= e)
        makespan += c
    return {"sum_of_cost": sum_of_cost, "makespan": makespan}


def run(config: dict, *, out_dir: Optional[str] = None, device: int = 0) -> List[dict]:
    """main(config) of eval.jl:98-174: generate instances, run every solver entry on every
    instance, validate, write result.csv (+ config.yaml) and return the rows."""
    import yaml
    target = config["instance"]["_target_"].split("gen_random_instance_")[-1]
    State = _STATE_BY_NAME[target]
    params = {k: v for k, v in config["instance"].items() if k != "_target_"}
    seed_offset = int(config.get("seed_offset", 0))
    num_instances = int(config.get("num_instances", 3))
    time_limit = config.get("time_limit", 10)
    rng = np.random.default_rng(seed_offset)
    instances = [gen_random_instance(State, rng, device=device, **params) for _ in range(num_instances)]
    rows = []
    for k, (init, goal, obstacles, ins_params) in enumerate(instances, start=1):
        tag = init[0]
        connect = CL.gen_connect(tag, obstacles, *ins_params, device=device)
        collide = CL.gen_collide(tag, *ins_params, device=device)
        check_goal = CL.gen_check_goal(goal)
        for l, info in enumerate(config["solvers"], start=1):
            name = info["_target_"]
            kw = {key: val for key, val in info.items() if key != "_target_"}
            solution, t_plan = None, 0.0
            if name.split(".")[-1] == "SSSP":
                t0 = time.perf_counter()
                solution, _ = SSSP(init, goal, connect, collide, check_goal, TIME_LIMIT=time_limit,
                                   seed=seed_offset, **kw)
                t_plan = time.perf_counter() - t0
                if not validate(init, connect, collide, check_goal, solution):
                    raise RuntimeError(f"{name} yields invalid solution for instance-{k}, "
                                       f"seed={seed_offset}: {validate.last_failure}")
            cost = get_solution_cost(connect.ctx, solution)
            # refinement (eval.jl:62-66): temporal plan graph + greedy re-timing
            t0 = time.perf_counter()
            refined = smoothing(solution, connect, collide)
            t_ref = time.perf_counter() - t0 if solution is not None else 0.0
            cost_ref = None if refined is None else refined[2]
            rows.append(dict(instance=k, N=len(init), num_obs=len(obstacles), solver=name, solver_index=l,
                             solved=solution is not None, elapsed_planning=t_plan,
                             elapsed_refinement=t_ref, elapsed_total=t_plan + t_ref,
                             sum_of_cost_original=0 if cost is None else cost["sum_of_cost"],
                             makespan_original=0 if cost is None else cost["makespan"],
                             sum_of_cost_refined=0 if cost_ref is None else cost_ref["sum_of_cost"],
                             makespan_refined=0 if cost_ref is None else cost_ref["makespan"]))
    if out_dir is not None:
        date_str = datetime.datetime.now().isoformat()
        root = os.path.join(out_dir, date_str)
        os.makedirs(root, exist_ok=True)
        with open(os.path.join(root, "config.yaml"), "w") as f:
            yaml.safe_dump({**config, "date": date_str}, f)
        with open(os.path.join(root, "result.csv"), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=RESULT_COLUMNS)
            w.writeheader()
            for r in rows:
                w.writerow({**r, "solved": "true" if r["solved"] else "false"})   # Julia Bool in CSV.jl
        run.last_dir = root
    return rows


run.last_dir = None


def planning(req: dict, device: int = 0) -> dict:
    """One request of the planning server (scripts/server.jl:84-122) -> its reply as a dict in the
    reference's JSON format.  req = {"field": {"x", "y", "size"}, "agents": [{"x_init", "y_init",
    "x_goal", "y_goal", "rad"}, ...], "obstacles": [{"x", "y", "rad"}, ...] | null,
    "solver": {"_target_": "MRMP.Solvers.SSSP", <keyword arguments>}}; coordinates are scaled into
    the unit square by the field (server.jl:20-27) and the instructions scaled back (:29-52)."""
    fx, fy, fs = float(req["field"]["x"]), float(req["field"]["y"]), float(req["field"]["size"])
    P = CL.StatePoint2D
    init = [P((float(a["x_init"]) - fx) / fs, (float(a["y_init"]) - fy) / fs) for a in req["agents"]]
    goal = [P((float(a["x_goal"]) - fx) / fs, (float(a["y_goal"]) - fy) / fs) for a in req["agents"]]
    obstacles = np.array([[(float(o["x"]) - fx) / fs, (float(o["y"]) - fy) / fs, float(o["rad"]) / fs]
                          for o in (req.get("obstacles") or [])]).reshape(-1, 3)
    rads = [float(a["rad"]) / fs for a in req["agents"]]
    connect = CL.gen_connect(init[0], obstacles, rads, device=device)
    collide = CL.gen_collide(init[0], rads, device=device)
    check_goal = CL.gen_check_goal(goal)
    if not is_valid_instance(init, goal, connect, collide):            # server.jl:100-101
        return {"status": "failure"}
    solver = req["solver"]["_target_"].split(".")[-1]
    params = {k: v for k, v in req["solver"].items() if k != "_target_"}
    if solver != "SSSP":
        raise ValueError(f"solver {req['solver']['_target_']} is not built on this path")
    solution, _ = SSSP(init, goal, connect, collide, check_goal, **params)
    if solution is None:                                               # server.jl:110
        return {"status": "failure"}
    TPG, solution, cost = smoothing(solution, connect, collide)
    return {"status": "success", "instructions": get_instructions(TPG, fx, fy, fs)}


def main(argv: Sequence[str]) -> None:
    """python -m sssp_b200.harness config.yaml [key.sub=value ...] (eval.jl:177-229)"""
    config = load_config(argv[0], argv[1:])
    out = config.get("root", os.path.join("data", "exp"))
    rows = run(config, out_dir=out)
    solved = sum(r["solved"] for r in rows)
    print(f"{solved}/{len(rows)} tasks solved; result saved in {run.last_dir}")


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
