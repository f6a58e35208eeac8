"""Generates tests/golden/sssp_*.json: seeded instances (the gen_random_instance procedure of
/root/reference/src/utils.jl:160-245 without its wall-clock retry, driven by numpy's PCG64
because Julia's RNG stream cannot be reproduced) and the ORACLE's SSSP result on them
(oracle/sssp_ref.py).  Run from the repo root:  python tests/golden/make_sssp_fixtures.py
The GPU tests compare the CUDA path with these files and with the live oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from oracle import sssp_ref as R        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


NPOS = {O.POINT2D: 2, O.POINT3D: 3, O.ARM22: 0, O.LINE2D: 2, O.SNAKE2D: 2, O.ARM33: 0, O.CAPSULE3D: 3}


def gen_instance(model, N, num_obs, rad, rad_obs, seed, base_box=(0.25, 0.75), axis=(0.08, 0.12)):
    """utils.jl:204-245 / :160-201: obstacles, radii, then per agent a start and a goal that are
    free (connect(q,i)) and do not collide with the starts / goals chosen so far."""
    rng = np.random.default_rng(seed)
    d, wd = O._D[model], O._W[model]
    obs = np.array([list(rng.random(wd)) + [rng.uniform(*rad_obs)] for _ in range(num_obs)]).reshape(-1, wd + 1)
    rads = np.array([rng.uniform(*rad) for _ in range(N)])
    base = rng.uniform(*base_box, size=(N, wd)) if model in (O.ARM22, O.ARM33) else None
    aux = rng.uniform(*axis, size=N) if model == O.CAPSULE3D else None
    orc = O.Oracle(model, rads, obs, base=base, aux=aux)

    def sample():   # gen_uniform_sampling: positions rand(), angles rand() * 2pi
        q = rng.random(d)
        q[NPOS[model]:] *= 2 * np.pi
        return q

    init, goal = [], []
    for i in range(N):
        for chosen in (init, goal):
            for _ in range(100000):
                q = sample()
                if orc.connect_point(q, i) and all(not orc.collide_static(q, e, i, j)
                                                   for j, e in enumerate(chosen)):
                    chosen.append(q)
                    break
            else:
                raise RuntimeError("instance generation stalled")
    return dict(model=int(model), rads=rads.tolist(), obstacles=obs.tolist(),
                base=None if base is None else base.tolist(),
                aux=None if aux is None else aux.tolist(),
                config_init=[q.tolist() for q in init], config_goal=[q.tolist() for q in goal])


CASES = {
    # BASELINE.json configs[0]: README minimum example (readme.md:51-57), SSSP defaults
    "sssp_point2d_readme_seed46": dict(model=O.POINT2D, N=5, num_obs=3, rad=(0.1, 0.1),
                                       rad_obs=(0.1, 0.1), seed=46, params=dict()),
    # configs[1]: point2d_many.yaml (epsilon 0.2, init_min_dist_thread 0.05), N = 10
    "sssp_point2d_many_n10_seed0": dict(model=O.POINT2D, N=10, num_obs=10,
                                        rad=(0.015625, 0.0234375), rad_obs=(0.025, 0.04), seed=0,
                                        params=dict(epsilon=0.2, init_min_dist_thread=0.05)),
    # configs[1] at the sizes north_star names (N = 30 and the target N = 50; 472 / 7811 expansions)
    "sssp_point2d_many_n30_seed0": dict(model=O.POINT2D, N=30, num_obs=10,
                                        rad=(0.015625, 0.0234375), rad_obs=(0.025, 0.04), seed=0,
                                        params=dict(epsilon=0.2, init_min_dist_thread=0.05)),
    "sssp_point2d_many_n50_seed0": dict(model=O.POINT2D, N=50, num_obs=10,
                                        rad=(0.015625, 0.0234375), rad_obs=(0.025, 0.04), seed=0,
                                        params=dict(epsilon=0.2, init_min_dist_thread=0.05)),
    # configs[2]: point3d.yaml solver parameters, radii reduced so that N agents fit
    "sssp_point3d_n6_seed1": dict(model=O.POINT3D, N=6, num_obs=6, rad=(0.05, 0.05),
                                  rad_obs=(0.05, 0.1), seed=1,
                                  params=dict(init_min_dist_thread=0.2, steering_depth=2,
                                              epsilon=0.5, num_vertex_expansion=30)),
    # per-model eval config of point2d (no_fast_collision_check, point2d.yaml:41-49)
    "sssp_point2d_slow_collide_seed3": dict(model=O.POINT2D, N=4, num_obs=5, rad=(0.05, 0.08),
                                            rad_obs=(0.05, 0.1), seed=3,
                                            params=dict(steering_depth=4, epsilon=4.0,
                                                        num_vertex_expansion=20,
                                                        no_fast_collision_check=True)),
    # configs[3]: arm22.yaml solver parameters (steering_depth 4, num_vertex_expansion 5)
    "sssp_arm22_n3_seed2": dict(model=O.ARM22, N=3, num_obs=1, rad=(0.12, 0.16),
                                rad_obs=(0.05, 0.08), seed=2,
                                params=dict(steering_depth=4, num_vertex_expansion=5)),
    # the other discretised models with their eval-config solver parameters
    # (scripts/config/eval/line2d.yaml, capsule3d.yaml), bodies sized so that N agents fit
    "sssp_line2d_n4_seed5": dict(model=O.LINE2D, N=4, num_obs=3, rad=(0.06, 0.1),
                                 rad_obs=(0.05, 0.08), seed=5,
                                 params=dict(steering_depth=2, num_vertex_expansion=10)),
    "sssp_capsule3d_n3_seed7": dict(model=O.CAPSULE3D, N=3, num_obs=3, rad=(0.03, 0.05),
                                    rad_obs=(0.05, 0.1), seed=7,
                                    params=dict(steering_depth=2, num_vertex_expansion=10,
                                                init_min_dist_thread=0.2)),
}


def main():
    only = set(sys.argv[1:])     # optional: names of the cases to (re)generate
    for name, c in CASES.items():
        if only and name not in only:
            continue
        inst = gen_instance(c["model"], c["N"], c["num_obs"], c["rad"], c["rad_obs"], c["seed"])
        orc = O.Oracle(c["model"], inst["rads"], np.array(inst["obstacles"]),
                       base=None if inst["base"] is None else np.array(inst["base"]),
                       aux=None if inst["aux"] is None else np.array(inst["aux"]))
        sol, rms, st = R.SSSP(orc, inst["config_init"], inst["config_goal"], seed=c["seed"],
                              **c["params"])
        out = dict(instance=inst, params=c["params"], seed=c["seed"], solution=sol,
                   expansions=st["expansions"],
                   roadmaps=[dict(nodes=[[float(x).hex() for x in v.q] for v in rm],
                                  neighbors=[list(map(int, v.neighbors)) for v in rm]) for rm in rms])
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, "solved" if sol else "FAILED", "steps", None if not sol else len(sol),
              "expansions", st["expansions"], "roadmap sizes", [len(r) for r in rms])


if __name__ == "__main__":
    main()
