# ref_bench.jl -- run the UNMODIFIED reference (Kei18/sssp, package MRMP) on a fixture written by
# tools/export_ref_fixture.py: (1) pins the CPU oracle of this repo against real Julia (neighbour
# lists of PRM!, collide verdicts, connect verdicts must agree bit for bit), (2) times the
# reference's own code on the same inputs (SURVEY.md 8(d): "if a julia binary is found").
#
#   julia --project=/path/to/sssp --threads=auto julia/ref_bench.jl fixture.json [reps]
#
# Prints ONE JSON line.  Not executable in the build image (no Julia there); bench.py calls it only
# when `julia` is on PATH and MRMP_REFERENCE_DIR points at a checkout of the reference.
#
# Only the reference's public API is used: gen_connect (point2d.jl:42-69), gen_collide
# (point.jl:5-36), Solvers.PRM! on a pre-filled roadmap (prm.jl:178-213: with num_vertices vertices present
# the sampling loop is skipped and only the neighbour pass runs).

import MRMP
import MRMP: StatePoint2D, CircleObstacle2D, Node, gen_connect, gen_collide
import JSON

function main()
    fx = JSON.parsefile(ARGS[1])
    reps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 3
    N, nv = fx["n_agents"], fx["num_vertices"]
    rad = Float64(fx["rad"])
    rads = Vector{Float64}(fx["rads"])
    obstacles = [CircleObstacle2D(Float64(o[1]), Float64(o[2]), Float64(o[3])) for o in fx["obstacles"]]
    q0 = StatePoint2D(0.0, 0.0)
    connect = gen_connect(q0, obstacles, rads)
    collide = gen_collide(q0, rads)
    V = [[StatePoint2D(Float64(v[1]), Float64(v[2])) for v in fx["vertices"][i]] for i = 1:N]

    # ---- (A) neighbour pass of PRM! on the fixture's vertices (agents are independent, prm.jl:265-275)
    function build()
        map(1:N) do i
            roadmap = [Node{StatePoint2D}(V[i][k], k, Int64[]) for k = 1:nv]
            MRMP.Solvers.PRM!(roadmap, (args...) -> connect(args..., i), nv, rad)
        end
    end
    roadmaps = build()
    t_build = minimum([(@elapsed build()) for _ = 1:reps])
    nnz = sum(length(v.neighbors) for rm in roadmaps for v in rm)
    bad_rows = 0
    for i = 1:N, k = 1:nv
        want = [Int64(x) + 1 for x in fx["neighbors"][i][k]]
        bad_rows += roadmaps[i][k].neighbors == want ? 0 : 1
    end

    # ---- (B) collide(q_from, q_to, p_from, p_to, i, j) on the table sample, the order PP asks it
    tb = fx["table"]
    ra, rf, rt = tb["row_agent"], tb["row_vfrom"], tb["row_vto"]
    ca, cf, ct = tb["col_agent"], tb["col_vfrom"], tb["col_vto"]
    n_rows, n_cols = length(ra), length(ca)
    function table()
        out = fill(false, n_rows, n_cols)   # Matrix{Bool}: a BitMatrix is not safe to fill from threads
        Threads.@threads for r = 1:n_rows
            i = ra[r] + 1
            a, b = V[i][rf[r]+1], V[i][rt[r]+1]
            for c = 1:n_cols
                j = ca[c] + 1
                j == i && continue
                out[r, c] = collide(a, b, V[j][cf[c]+1], V[j][ct[c]+1], i, j)
            end
        end
        out
    end
    got = table()
    t_table = minimum([(@elapsed table()) for _ = 1:reps])
    bad_pairs = 0
    for r = 1:n_rows
        s = tb["verdict"][r]
        for c = 1:n_cols
            bad_pairs += (got[r, c] == (s[c] == '1')) ? 0 : 1
        end
    end
    n_pairs = sum(ca[c] != ra[r] for r = 1:n_rows, c = 1:n_cols)

    # ---- (C) connect(q, i)
    pt = fx["points"]
    bad_points = 0
    for k = 1:length(pt["agent"])
        q = StatePoint2D(Float64(pt["q"][k][1]), Float64(pt["q"][k][2]))
        bad_points += (connect(q, pt["agent"][k] + 1) == (pt["connect"][k] == 1)) ? 0 : 1
    end

    println(JSON.json(Dict(
        "impl" => "reference-julia", "julia" => string(VERSION), "threads" => Threads.nthreads(),
        "n_agents" => N, "num_vertices" => nv, "roadmap_nnz" => nnz,
        "roadmap_build_ms" => 1e3 * t_build,            # single-threaded, like a reference PRMs call
        "table_pairs" => n_pairs, "table_s" => t_table,
        "collide_pair_checks_per_s" => n_pairs / t_table,
        "mismatching_neighbor_rows_vs_oracle" => bad_rows,
        "mismatching_collide_pairs_vs_oracle" => bad_pairs,
        "mismatching_connect_points_vs_oracle" => bad_points,
        "oracle_pinned" => bad_rows == 0 && bad_pairs == 0 && bad_points == 0,
    )))
end

main()
