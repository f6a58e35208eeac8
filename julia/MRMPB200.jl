# MRMPB200.jl -- ccall glue: rebuilds the reference's closure API on libmrmp_b200.so.
#
# NOT EXECUTED in this repository's CI: there is no Julia runtime in the build image or on
# the GPU box (see DESIGN.md).  It documents, line by line, the binding a maintainer of
# Kei18/sssp would add; the C entry points it calls are exercised through the identical
# ctypes binding in sssp_b200/lib.py.
#
# Usage inside MRMP.jl (replacing src/models/point2d.jl:42-69 and src/models/point.jl:5-36):
#   connect = MRMPB200.gen_connect(StatePoint2D(0, 0), obstacles, rads)
#   collide = MRMPB200.gen_collide(StatePoint2D(0, 0), rads)
# Both return multi-method closures with the reference's signatures (1-based agent ids); the
# same constructors serve StatePoint3D and StateArm22 (ins_params = (positions, rads)).
# SsspRoadmaps / expand_and_filter! (bottom of the file) are the batched node expansion.
module MRMPB200

import ..MRMP: AbsState, Node, StatePoint2D, StatePoint3D, StateArm22, StateLine2D, StateSnake2D,
               StateArm33, StateCapsule3D, CircleObstacle2D, CircleObstacle3D, STEP_DIST,
               SAFETY_DIST_LINE, STEP_DIST_SNAKE2D, STEP_DIST_CAPSULE3D

const LIB = get(ENV, "MRMP_B200_LIB", "libmrmp_b200.so")
const MODEL = Dict(StatePoint2D => 0, StatePoint3D => 1, StateArm22 => 2, StateLine2D => 3,
                   StateSnake2D => 4, StateArm33 => 5, StateCapsule3D => 6)
# per-model default of the step_dist keyword (snake2d.jl:3, capsule3d.jl:3)
default_step(::Type{StateSnake2D}) = STEP_DIST_SNAKE2D
default_step(::Type{StateCapsule3D}) = STEP_DIST_CAPSULE3D
default_step(::Type) = STEP_DIST

# Objects created on a ctx use it in their own destroy call and Julia runs finalizers in no
# particular order: the ctx is destroyed when its own finalizer has run AND its last dependent
# (SsspRoadmaps, ...) is gone.
mutable struct Ctx
    h::Ptr{Cvoid}
    deps::Int
    dead::Bool
    function Ctx(h)
        c = new(h, 0, false)
        finalizer(c) do x
            x.dead = true
            x.deps == 0 && destroy_ctx!(x)
        end
        return c
    end
end
function destroy_ctx!(c::Ctx)
    c.h == C_NULL && return
    ccall((:mrmp_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h)
    c.h = C_NULL
end
adopt!(c::Ctx) = (c.deps += 1; c)
function release!(c::Ctx)
    c.deps -= 1
    c.dead && c.deps == 0 && destroy_ctx!(c)
end

function check(rc::Cint)
    rc == 0 && return
    error("mrmp_b200: " * unsafe_string(ccall((:mrmp_last_error, LIB), Cstring, ())))
end

# isbits structs are C-layout: Vector{StatePoint2D} == double[n][2], Vector{CircleObstacle2D}
# == double[n][3]  (src/obstacle.jl:7-19), so they are passed without conversion.
function make_ctx(q::State, obstacles, rads::Vector{Float64};
                  positions = nothing, axises = nothing, step_dist = default_step(State),
                  safety_dist = SAFETY_DIST_LINE, max_dist = nothing,
                  device = 0) where {State<:AbsState}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    # positions::Vector{Vector{Float64}} (arm22.jl:32) -> one flat rooted array; every array whose
    # raw pointer crosses the ccall is kept alive by GC.@preserve (the library copies them)
    flat = isnothing(positions) ? Float64[] : reduce(vcat, positions)
    ax = isnothing(axises) ? Float64[] : axises                   # capsule3d.jl:65
    GC.@preserve flat ax obstacles begin
        base = isnothing(positions) ? Ptr{Float64}(C_NULL) : pointer(flat)
        aux = isnothing(axises) ? Ptr{Float64}(C_NULL) : pointer(ax)
        check(ccall((:mrmp_ctx_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
                     Ptr{Cvoid}, Float64, Float64, Float64, Cint),
                    h, MODEL[State], length(rads), rads, base, aux, length(obstacles),
                    isempty(obstacles) ? C_NULL : pointer(obstacles), step_dist, safety_dist,
                    isnothing(max_dist) ? NaN : max_dist, device))
    end
    return Ctx(h[])
end

"""gen_connect(q, obstacles, rads) -- src/models/point2d.jl:42-69, point3d.jl:47-80"""
function gen_connect(q::State, obstacles, ins_params...; kwargs...) where {State<:AbsState}
    # ins_params: (rads,) | (positions, rads) for arm22 / arm33 | (rads, axises) for capsule3d
    capsule = State == StateCapsule3D
    rads = capsule ? ins_params[1] : ins_params[end]
    positions = (!capsule && length(ins_params) == 2) ? ins_params[1] : nothing   # arm22.jl:29-37
    ctx = make_ctx(q, obstacles, rads; positions = positions,
                   axises = capsule ? ins_params[2] : nothing, kwargs...)
    out = Ref{UInt8}(0)

    f(q::State, i::Int64)::Bool = begin                       # point2d.jl:49-57
        check(ccall((:mrmp_connect_point, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ref{State}, Ref{UInt8}), ctx.h, i - 1, q, out))
        return out[] != 0
    end
    f(q_from::State, q_to::State, i::Int64)::Bool = begin     # point2d.jl:59-66
        check(ccall((:mrmp_connect_edge, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ref{State}, Ref{State}, Ref{UInt8}),
                    ctx.h, i - 1, q_from, q_to, out))
        return out[] != 0
    end
    # new batched method: one launch for a whole edge list
    f(q_from::Vector{State}, q_to::Vector{State}, agents::Vector{Int32})::Vector{Bool} = begin
        res = Vector{Bool}(undef, length(agents))
        check(ccall((:mrmp_connect_edges, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{State}, Ptr{State}, Ptr{Bool}),
                    ctx.h, length(agents), agents .- Int32(1), q_from, q_to, res))
        return res
    end
    return f
end

"""gen_collide(q, rads) -- src/models/point.jl:5-36, arm22.jl:90-212"""
function gen_collide(q::State, ins_params...; kwargs...) where {State<:AbsState}
    capsule = State == StateCapsule3D
    rads = capsule ? ins_params[1] : ins_params[end]
    positions = (!capsule && length(ins_params) == 2) ? ins_params[1] : nothing
    ws3 = State in (StatePoint3D, StateArm33, StateCapsule3D)
    ctx = make_ctx(q, ws3 ? CircleObstacle3D[] : CircleObstacle2D[], rads;
                   positions = positions, axises = capsule ? ins_params[2] : nothing, kwargs...)
    N = length(rads)
    out = Ref{UInt8}(0)

    f(a::State, b::State, c::State, d::State, i::Int64, j::Int64)::Bool = begin   # point.jl:9-12
        check(ccall((:mrmp_collide_pair, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Ref{State}, Ref{State}, Ref{State}, Ref{State},
                     Ref{UInt8}), ctx.h, i - 1, j - 1, a, b, c, d, out))
        return out[] != 0
    end
    f(q_i::State, q_j::State, i::Int64, j::Int64)::Bool = begin                   # point.jl:14-16
        ii, jj = Int32[i - 1], Int32[j - 1]
        check(ccall((:mrmp_collide_static_pairs, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ref{State}, Ref{State},
                     Ref{UInt8}), ctx.h, 1, ii, jj, q_i, q_j, out))
        return out[] != 0
    end
    f(Q_from::Vector{Node{State}}, Q_to::Vector{Node{State}})::Bool = begin       # point.jl:18-25
        # Vector{Node} is an array of pointers (src/MRMP.jl:73-77): gather the states flat,
        # exactly what the reference itself does at arm22.jl:207-209
        qf, qt = map(v -> v.q, Q_from), map(v -> v.q, Q_to)
        check(ccall((:mrmp_collide_configs, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{State}, Ptr{State}, Ref{UInt8}, Ptr{Int64}),
                    ctx.h, 1, qf, qt, out, C_NULL))
        return out[] != 0
    end
    f(Q::Vector{Node{State}}, q_to::State, i::Int64)::Bool = begin                # point.jl:27-33
        qs = map(v -> v.q, Q)
        check(ccall((:mrmp_collide_config_moves, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{State}, Int64, Ref{State}, Ref{UInt8}),
                    ctx.h, i - 1, qs, 1, q_to, out))
        return out[] != 0
    end
    # new batched method used by SSSP's successor filter (sssp.jl:196-223): all candidate
    # successors p of agent i against the static others in one launch
    f(Q::Vector{Node{State}}, q_tos::Vector{State}, i::Int64)::Vector{Bool} = begin
        qs = map(v -> v.q, Q)
        res = Vector{Bool}(undef, length(q_tos))
        check(ccall((:mrmp_collide_config_moves, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{State}, Int64, Ptr{State}, Ptr{Bool}),
                    ctx.h, i - 1, qs, length(q_tos), q_tos, res))
        return res
    end
    return f
end

"""PRMs(config_init, config_goal, ctx, num_vertices; rad) -- src/solvers/prm.jl:251-276.
Sampling + neighbour pass for all agents on the GPU; returns Vector{Vector{Node{State}}}."""
function PRMs(config_init::Vector{State}, config_goal::Vector{State}, ctx::Ctx,
              num_vertices::Int64 = 100; rad::Union{Nothing,Float64} = nothing,
              seed::UInt64 = UInt64(0)) where {State<:AbsState}
    N = length(config_init)
    d = sizeof(State) ÷ 8
    nodes = Vector{State}(undef, N * num_vertices)
    row_ptr = Vector{Int32}(undef, N * num_vertices + 1)
    cap = N * num_vertices * 64
    col = Vector{Int32}(undef, cap)
    nnz = Ref{Int64}(0)
    rads = fill(isnothing(rad) ? NaN : rad, N)
    rc = ccall((:mrmp_prms, LIB), Cint,
               (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, UInt64, Ptr{State}, Ptr{State},
                Ptr{State}, Ptr{Int32}, Ptr{Int32}, Int64, Ref{Int64}),
               ctx.h, 0, N, num_vertices, rads, seed, config_init, config_goal, nodes, row_ptr,
               col, cap, nnz)
    if rc == -3          # MRMP_ERR_CAPACITY: retry with the reported size
        resize!(col, nnz[]); cap = nnz[]
        rc = ccall((:mrmp_prms, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, UInt64, Ptr{State}, Ptr{State},
                    Ptr{State}, Ptr{Int32}, Ptr{Int32}, Int64, Ref{Int64}),
                   ctx.h, 0, N, num_vertices, rads, seed, config_init, config_goal, nodes,
                   row_ptr, col, cap, nnz)
    end
    check(rc)
    return [[Node{State}(nodes[(a - 1) * num_vertices + v], v,
                         Int64.(col[row_ptr[(a - 1) * num_vertices + v] + 1:
                                    row_ptr[(a - 1) * num_vertices + v + 1]]) .+ 1)
             for v = 1:num_vertices] for a = 1:N]
end

# ---- SSSP: device-resident growing roadmaps + one call per popped search node -------------
# Replaces, inside src/solvers/sssp.jl:181-223, the call `expand!(...)` (:181-189) and the
# per-successor `collide(S.Q, p.q, i)` (:206) by a single library call.  The search loop
# itself (OPEN, VISITED, h_func, backtrack) is unchanged Julia.
mutable struct SsspRoadmaps
    h::Ptr{Cvoid}
    ctx::Ctx
    function SsspRoadmaps(ctx::Ctx, nv_cap::Int = 256)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mrmp_sssp_create, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}),
                    ctx.h, nv_cap, h))
        s = new(h[], adopt!(ctx))
        finalizer(s) do x
            ccall((:mrmp_sssp_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h)
            release!(x.ctx)
        end
        return s
    end
end

"""upload agent i's initial roadmap (output of gen_RRT_connect_roadmap, sssp.jl:325-401)"""
function set_roadmap!(s::SsspRoadmaps, i::Int64, roadmap::Vector{Node{State}}) where {State}
    qs = map(v -> v.q, roadmap)
    check(ccall((:mrmp_sssp_set_roadmap, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{State}),
                s.h, i - 1, length(qs), qs))
end

"""
    expand_and_filter!(s, i, v, S_Q, roadmap_i; kwargs...) -> Vector{Tuple{Int64,Bool}}

expand! (sssp.jl:235-267) for agent i at vertex v plus the successor filter (:196-223):
appends the accepted vertices and their edges to `roadmap_i` in the reference's order and
returns `(p_id, collides)` for every `p_id in vcat(v.neighbors, v.id)`.
"""
function expand_and_filter!(s::SsspRoadmaps, i::Int64, v::Node{State}, S_Q::Vector{Node{State}},
                            roadmap::Vector{Node{State}}; num_vertex_expansion::Int64 = 10,
                            steering_depth::Int64 = 2, epsilon::Union{Nothing,Float64} = nothing,
                            min_dist_thread::Float64 = 0.1, seed::UInt64 = UInt64(0),
                            t0::UInt64 = UInt64(0), slow::Bool = false) where {State<:AbsState}
    K = num_vertex_expansion
    nv0 = length(roadmap)
    words = (nv0 + K + 31) ÷ 32
    q_new = Vector{State}(undef, K)
    out_bits, in_bits = zeros(UInt32, words, K), zeros(UInt32, words, K)   # column r = row r in C
    old = Int32.(v.neighbors .- 1)
    succ, hit = Vector{Int32}(undef, length(old) + K + 1), Vector{UInt8}(undef, length(old) + K + 1)
    n_new, n_succ, w = Ref{Int32}(0), Ref{Int32}(0), Ref{Int32}(0)
    Q_ids = Int32[u.id - 1 for u in S_Q]
    check(ccall((:mrmp_sssp_expand, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}, UInt64, UInt64, Cint, Float64, Float64,
                 Ptr{Int32}, Cint, Ptr{Int32}, Cint, Ref{Int32}, Ptr{State}, Ref{Int32},
                 Ptr{UInt32}, Ptr{UInt32}, Int64, Ref{Int32}, Ptr{Int32}, Ptr{UInt8}),
                s.h, i - 1, v.id - 1, K, C_NULL, seed, t0, steering_depth,
                isnothing(epsilon) ? NaN : epsilon, min_dist_thread, Q_ids, length(old), old,
                slow ? 1 : 0, n_new, q_new, w, out_bits, in_bits, length(out_bits), n_succ, succ,
                hit))
    bit(B, r, t) = (B[(t >> 5) + 1, r] >> (t & 31)) & 0x1 == 0x1
    for r = 1:n_new[]                      # same push! order as sssp.jl:252-262
        id = nv0 + r
        u = Node{State}(q_new[r], id, [t + 1 for t = 0:id-2 if bit(out_bits, r, t)])
        push!(roadmap, u)
        foreach(t -> push!(roadmap[t+1].neighbors, id), [t for t = 0:id-1 if bit(in_bits, r, t)])
    end
    return [(Int64(succ[k]) + 1, hit[k] != 0) for k = 1:n_succ[]]
end

# ---- PP (pp.jl:140-260): conflict tables for the agent that is planned next ------------------
"""
    pp_conflict_tables(ctx, roadmaps, i, paths, max_makespan) -> (blocked, stay_hits, planned)

Everything `invalid` (pp.jl:171-193) and `check_goal_i` (pp.jl:195-213) ask about agent `i`,
in two cross-product launches instead of one `collide` call per generated search node.
Rows are agent i's moves in expansion order `vcat(v.neighbors, v.id)` (solvers/utils.jl:107),
columns `(paths[j][min(t-1,l)], paths[j][min(t,l)])` for `t = 2:max_makespan`, j over the
planned agents.  `blocked[move, t-1]` is the OR over j (taken on the device, group_size = K);
`stay_hits[v, (t-2)*K + k]` is the verdict of the stay move of vertex v against planned[k] at t.
"""
function pp_conflict_tables(ctx::Ctx, roadmaps::Vector{Vector{Node{State}}}, i::Int64,
                            paths::Vector{Vector{Node{State}}},
                            max_makespan::Int64) where {State<:AbsState}
    rm = roadmaps[i]
    planned = [j for j = 1:length(roadmaps) if j != i && !isempty(paths[j])]
    K, T = length(planned), max_makespan - 1
    rf = State[v.q for v in rm for _ in 1:(length(v.neighbors) + 1)]
    rt = State[rm[k].q for v in rm for k in vcat(v.neighbors, v.id)]
    ca = Int32[j - 1 for _ = 2:max_makespan for j in planned]
    cf = State[paths[j][min(t - 1, length(paths[j]))].q for t = 2:max_makespan for j in planned]
    ct = State[paths[j][min(t, length(paths[j]))].q for t = 2:max_makespan for j in planned]
    ra = fill(Int32(i - 1), length(rf))
    wpr = (T + 31) ÷ 32
    bits = zeros(UInt32, wpr, length(rf))                  # column r = row r in C
    check(ccall((:mrmp_collide_cross, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{State}, Ptr{State}, Int64, Ptr{Int32},
                 Ptr{State}, Ptr{State}, Cint, Ptr{UInt32}),
                ctx.h, length(rf), ra, rf, rt, length(ca), ca, cf, ct, K, bits))
    blocked = [(bits[(c >> 5) + 1, r] >> (c & 31)) & 0x1 == 0x1 for r = 1:length(rf), c = 0:T-1]
    sq = State[v.q for v in rm]
    swpr = (T * K + 31) ÷ 32
    sbits = zeros(UInt32, swpr, length(rm))
    check(ccall((:mrmp_collide_cross, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{State}, Ptr{State}, Int64, Ptr{Int32},
                 Ptr{State}, Ptr{State}, Cint, Ptr{UInt32}),
                ctx.h, length(rm), fill(Int32(i - 1), length(rm)), sq, sq, length(ca), ca, cf, ct,
                0, sbits))
    stay_hits = [(sbits[(c >> 5) + 1, v] >> (c & 31)) & 0x1 == 0x1 for v = 1:length(rm), c = 0:T*K-1]
    return blocked, stay_hits, planned
end

# ---- CBS (cbs.jl:336-347) and the smoother's TPG scan (smoother.jl:78-106) ---------------------
const CROSS_COUNT, CROSS_FIRST, CROSS_FLAG_CBS = Cint(2), Cint(3), Cint(1)

"""
    cbs_collision_counts(ctx, roadmap_i, i, paths, T) -> counts[move, t-1]

The soft conflict count of `g_func_i` (cbs.jl:336-347) for every candidate move of agent `i`
(rows, `vcat(v.neighbors, v.id)` per vertex) and every time step `t = 2:T` in one launch:
`count(j -> collide(q_j_to, q_i_to, j, i) || collide(q_j_from, q_j_to, q_i_from, q_i_to, j, i))`
over the other agents' path edges.  The low-level search adds `counts[move, t-1] * collision_weight`.
"""
function cbs_collision_counts(ctx::Ctx, rm::Vector{Node{State}}, i::Int64,
                              paths::Vector{Vector{Node{State}}}, T::Int64) where {State<:AbsState}
    others = [j for j = 1:length(paths) if j != i]
    rf = State[v.q for v in rm for _ in 1:(length(v.neighbors) + 1)]
    rt = State[rm[k].q for v in rm for k in vcat(v.neighbors, v.id)]
    ca = Int32[j - 1 for _ = 2:T for j in others]
    cf = State[paths[j][min(t - 1, length(paths[j]))].q for t = 2:T for j in others]
    ct = State[paths[j][min(t, length(paths[j]))].q for t = 2:T for j in others]
    counts = zeros(Int32, T - 1, length(rf))                # column r = row r in C
    check(ccall((:mrmp_collide_cross_reduce, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{State}, Ptr{State}, Ptr{Int32}, Int64,
                 Ptr{Int32}, Ptr{State}, Ptr{State}, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Cint,
                 Ptr{Int32}),
                ctx.h, length(rf), fill(Int32(i - 1), length(rf)), rf, rt, C_NULL, length(ca), ca,
                cf, ct, C_NULL, C_NULL, length(others), T - 1, CROSS_COUNT, CROSS_FLAG_CBS, counts))
    return permutedims(counts)
end

"""
    tpg_first_hits(ctx, actions) -> first[action, agent]  (1-based index into `actions`, 0 = none)

Type-2 dependencies of `get_temporal_plan_graph` (smoother.jl:78-106): for every action and every
other agent the first LATER action of that agent that collides with it -- one launch instead of
the triple loop of `collide` calls.  `actions` = all actions ordered by agent, then by time.
"""
function tpg_first_hits(ctx::Ctx, actions::Vector{Action{State}}, N::Int64) where {State<:AbsState}
    ag = Int32[a.agent - 1 for a in actions]
    key = Int32[a.t for a in actions]
    qf = State[a.from.q for a in actions]
    qt = State[a.to.q for a in actions]
    first = fill(Int32(-1), N, length(actions))
    check(ccall((:mrmp_collide_cross_reduce, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{State}, Ptr{State}, Ptr{Int32}, Int64,
                 Ptr{Int32}, Ptr{State}, Ptr{State}, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Cint,
                 Ptr{Int32}),
                ctx.h, length(actions), ag, qf, qt, key, length(actions), ag, qf, qt, ag, key, 0, N,
                CROSS_FIRST, 0, first))
    return permutedims(first) .+ Int32(1)
end

# ---- pinned host arrays: async read-back / in-place conflict tables ---------------------------
"""
    with_pinned(f, arrays...)

Page-locks the given `Array`s for the duration of `f()` (mrmp_host_register): OR-reduced conflict
tables are then written by the kernel straight into the Julia array, and `read_async!` /
`read_wait` overlap the roadmap read-back with the next batch.
"""
function with_pinned(f, arrays...)
    foreach(a -> check(ccall((:mrmp_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), a, sizeof(a))),
            arrays)
    try
        return f()
    finally
        foreach(a -> ccall((:mrmp_host_unregister, LIB), Cint, (Ptr{Cvoid},), a), arrays)
    end
end

end # module
