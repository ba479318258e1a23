# MaviCUDA.jl — Julia glue that makes libmavi_cuda.so a drop-in `DeviceMode` of Mavi.jl.
#
# Usage (unchanged Mavi API, one new device type):
#
#     using Mavi, Mavi.Configs, MaviCUDA
#     system = System(state=..., space_cfg=..., dynamic_cfg=LenJonesCfg(sigma=1, epsilon=1),
#                     int_cfg=IntCfg(dt=0.001, chunks_cfg=ChunksCfg(900, 900), device=CUDADevice()))
#     Mavi.run_system(system, tf=1)          # one mavi_step(h, nsteps) + one download
#
# Everything below binds EXACTLY the entry points of include/mavi.h (the same ones the Python host mirror in
# mavi.jl_b200/ exercises through ctypes).  Julia is not installed in the build image, so this file is not executed by
# the test-suite; it is kept small and mechanical.  Reference seams it plugs into:
#   DeviceMode                  src/configs.jl:471-475
#   calc_forces!(system, chunks, device)   src/integration.jl:112,159,197,226
#   get_step_function           src/integration.jl:537-548, src/rings/integration.jl:545
#   System ctor                 src/systems.jl:73-114
module MaviCUDA

using StaticArrays
using Mavi
using Mavi.Configs
using Mavi.States
using Mavi.Systems
import Mavi.Integration: calc_forces!, newton_step!, szabo_step!, rtp_step!, get_step_function
import Mavi.RunSystem: run_system

export CUDADevice, sync_to_host!, upload_state!, device_energies, sync_rings_info!, sync_particle_neighbors!

const LIB = get(ENV, "MAVI_CUDA_LIB", "libmavi_cuda.so")
const MAVI_MAX_SPACES = 8

"`IntCfg(device=CUDADevice())` routes the per-step hot path to the GPU."
Base.@kwdef struct CUDADevice <: DeviceMode
    device::Int32 = 0
    rng_mode::Symbol = :philox          # :host_noise (caller passes the draws) | :philox (device RNG)
    seed::UInt64 = 24042001
    sync_every::Int = 1                  # download state every k calls of a per-step function (GUI / experiments)
    n_gpus::Int32 = 1                    # > 1: this ONE process drives devices device .. device+n_gpus-1 through the one
                                         # handle (MaviParams.n_gpus): x-slabs of cell columns inside the library, halo and
                                         # migration over NCCL/NVLink; the System is used exactly as with one GPU
end

# ---- flat PODs of include/mavi.h (isbits, same field order) -------------------------------------------------------
struct MaviLine
    p1::NTuple{2,Float64}
    p2::NTuple{2,Float64}
end

struct MaviSpace
    wall::Int32
    geom::Int32
    rect_bl::NTuple{2,Float64}
    rect_len::Float64
    rect_h::Float64
    circ_center::NTuple{2,Float64}
    circ_radius::Float64
    lines::Ptr{MaviLine}
    n_lines::Int32
    pot_kind::Int32
    pot::NTuple{4,Float64}
    pot_mode::Int32
    n_pot_types::Int32                  # PotentialVector (src/configs.jl:454-463): pot_types[t] for particles of type t
    pot_types::NTuple{16,Float64}       # [MAVI_MAX_POT_TYPES = 4][4], row-major
end

struct MaviRingsParams
    num_types::Int32
    n_max::Int32
    num_rings::Int64
    p0::Ptr{Float64}; relax_time::Ptr{Float64}; vo::Ptr{Float64}; mobility::Ptr{Float64}
    rot_diff::Ptr{Float64}; k_area::Ptr{Float64}; k_spring::Ptr{Float64}; l_spring::Ptr{Float64}
    num_particles::Ptr{Int32}
    interaction::Ptr{Float64}
    types::Ptr{Int32}
end

struct MaviParams
    struct_size::UInt32
    dtype::Int32
    n::Int64
    n_spaces::Int32
    _pad0::Int32
    spaces::NTuple{MAVI_MAX_SPACES,MaviSpace}
    grid_bl::NTuple{2,Float64}
    grid_len::Float64
    grid_h::Float64
    num_cols::Int32
    num_rows::Int32
    dynamics::Int32
    _pad1::Int32
    dyn::NTuple{8,Float64}
    particle_radius::Float64
    rings::Ptr{MaviRingsParams}
    dt::Float64
    rng_mode::Int32
    n_gpus::Int32
    seed::UInt64
    device::Int32
    flags::Int32
    stream::Ptr{Cvoid}
    rank::Int32
    world::Int32
    nccl_unique_id::Ptr{Cvoid}
    n_global::Int64
end

const WALL = Dict(RigidWalls => 0, PeriodicWalls => 1, SlipperyWalls => 2)
zero_space() = MaviSpace(0, 0, (0.0, 0.0), 0.0, 0.0, (0.0, 0.0), 0.0, C_NULL, 0, 0, (0.0, 0.0, 0.0, 0.0), 0, 0, ntuple(_ -> 0.0, 16))

function lower_space(w, g, keep)
    wall = w isa PotentialWalls ? Int32(3) : Int32(WALL[typeof(w)])
    pot_kind, pot, mode = Int32(0), (0.0, 0.0, 0.0, 0.0), Int32(2)
    npt, pts = Int32(0), zeros(Float64, 16)
    pot_of(p) = p isa LenJonesCfg ? (Int32(1), (Float64(p.sigma), Float64(p.epsilon), 0.0, 0.0)) :
                (Int32(0), (Float64(p.k_rep), Float64(p.k_atr), Float64(p.dist_eq), Float64(p.dist_max)))
    if w isa PotentialWalls
        p = w.potential
        if p isa Configs.PotentialVector                  # one potential per particle (ring) type, all of one kind
            length(p.vector) <= 4 || error("PotentialVector: at most 4 types on the device")
            npt = Int32(length(p.vector))
            for (t, q) in enumerate(p.vector)
                pot_kind, pq = pot_of(q)
                pts[4t-3:4t] .= pq
                t == 1 && (pot = pq)
            end
        else
            pot_kind, pot = pot_of(p)
        end
        mode = w.mode isa Configs.Outside ? Int32(0) : w.mode isa Configs.Inside ? Int32(1) : Int32(2)
    end
    pts = Tuple(pts)
    if g isa RectangleCfg
        return MaviSpace(wall, 0, Tuple(Float64.(g.bottom_left)), g.length, g.height, (0.0, 0.0), 0.0, C_NULL, 0, pot_kind, pot, mode, npt, pts)
    elseif g isa CircleCfg
        return MaviSpace(wall, 1, (0.0, 0.0), 0.0, 0.0, Tuple(Float64.(g.center)), g.radius, C_NULL, 0, pot_kind, pot, mode, npt, pts)
    else
        lines = [MaviLine(Tuple(Float64.(l.p1)), Tuple(Float64.(l.p2))) for l in g.lines]
        push!(keep, lines)
        return MaviSpace(wall, 2, (0.0, 0.0), 0.0, 0.0, (0.0, 0.0), 0.0, pointer(lines), length(lines), pot_kind, pot, mode, npt, pts)
    end
end

dyn_block(c::LenJonesCfg) = (Int32(0), (c.sigma, c.epsilon, 0, 0, 0, 0, 0, 0))
dyn_block(c::HarmTruncCfg) = (Int32(1), (c.k_rep, c.k_atr, c.dist_eq, c.dist_max, 0, 0, 0, 0))
dyn_block(c::SzaboCfg) = (Int32(2), (c.vo, c.mobility, c.relax_time, c.k_rep, c.k_adh, c.r_eq, c.r_max, c.rot_diff))
dyn_block(c::RunTumbleCfg) = (Int32(3), (c.vo, c.sigma, c.epsilon, c.tumble_rate, 0, 0, 0, 0))
dyn_block(::Mavi.Rings.Configs.RingsCfg) = (Int32(4), (0, 0, 0, 0, 0, 0, 0, 0))   # parameters travel in MaviRingsParams

"""
RingsCfg + RingsState -> MaviRingsParams (src/rings/configs.jl:95-107, src/rings/states.jl:74-124).  Scalars are
broadcast to one entry per ring type; the InteractionMatrix becomes [t1][t2][k_rep,k_atr,dist_eq,dist_max] (row-major for
the C side).  Every array is pushed to `keep` so that it outlives the `mavi_create` call.
"""
function lower_rings(cfg, state, keep)
    nt = cfg.num_types
    vecf(x) = (v = x isa Number ? fill(Float64(x), nt) : Float64.(x); push!(keep, v); v)
    np = state.num_particles isa Integer ? fill(Int32(state.num_particles), nt) : Int32.(state.num_particles)
    push!(keep, np)
    finder = cfg.interaction_finder
    inter = Float64[]
    for a in 1:nt, b in 1:nt
        ic = Mavi.Rings.Configs.get_interaction_cfg(a, b, finder)
        append!(inter, (ic.k_rep, ic.k_atr, ic.dist_eq, ic.dist_max))
    end
    push!(keep, inter)
    types = isnothing(state.types) ? Int32[] : Int32.(state.types)       # 1-based, as MaviRingsParams.types expects
    push!(keep, types)
    n_max, num_rings = size(state.rings_pos)
    rp = Ref(MaviRingsParams(nt, n_max, num_rings,
        pointer(vecf(cfg.p0)), pointer(vecf(cfg.relax_time)), pointer(vecf(cfg.vo)), pointer(vecf(cfg.mobility)),
        pointer(vecf(cfg.rot_diff)), pointer(vecf(cfg.k_area)), pointer(vecf(cfg.k_spring)), pointer(vecf(cfg.l_spring)),
        pointer(np), pointer(inter), isempty(types) ? C_NULL : pointer(types)))
    push!(keep, rp)
    return Base.unsafe_convert(Ptr{MaviRingsParams}, rp)
end

# ---- handle attached to a System (kept in a side table so that `System` itself is untouched) -----------------------
mutable struct DeviceState
    h::Ptr{Cvoid}
    keep::Vector{Any}
    calls::Int
end
const HANDLES = IdDict{Any,DeviceState}()

function check(ds::DeviceState, status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:mavi_last_error, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32), ds.h, buf, 512)
    error("libmavi_cuda status $status: " * unsafe_string(pointer(buf)))   # reference: throw(...) / BoundsError
end

second_array(s::SecondLawState) = s.vel
second_array(s::SelfPropelledState) = s.pol_angle
second_array(s::Mavi.Rings.States.RingsState) = s.pol          # one polarisation angle per ring

"Create the device context for `system` and upload its state (what the `System` ctor does on the CPU: force buffers, Chunks, first update_chunks!)."
function attach!(system::System)
    dev = system.int_cfg.device::CUDADevice
    keep = Any[]
    sc = system.space_cfg
    pairs = sc.wall_type isa ManyWalls ? collect(zip(sc.wall_type.list, sc.geometry_cfg.list)) : [(sc.wall_type, sc.geometry_cfg)]
    spaces = [lower_space(w, g, keep) for (w, g) in pairs]
    while length(spaces) < MAVI_MAX_SPACES; push!(spaces, zero_space()); end
    bbox = Configs.get_bounding_box(sc.geometry_cfg)
    cc = system.int_cfg.chunks_cfg
    kind, dyn = dyn_block(system.dynamic_cfg)      # unknown DynamicCfg -> MethodError -> caller keeps the CPU path
    rings_ptr = kind == 4 ? lower_rings(system.dynamic_cfg, system.state, keep) : Ptr{MaviRingsParams}(C_NULL)
    T = eltype(eltype(system.state.pos))            # Float64 (default) or Float32: the state's element type picks the build
    T in (Float64, Float32) || error("CUDADevice supports Float64 and Float32 states, got $T")
    params = Ref(MaviParams(sizeof(MaviParams), T === Float32 ? 1 : 0, length(system.state.pos), length(pairs), 0, Tuple(spaces),
        Tuple(Float64.(bbox.bottom_left)), bbox.length, bbox.height,
        isnothing(cc) ? 0 : cc.num_cols, isnothing(cc) ? 0 : cc.num_rows, kind, 0, Float64.(dyn),
        minimum(particle_radius(system.dynamic_cfg)), rings_ptr, system.int_cfg.dt,
        dev.rng_mode == :host_noise ? 0 : 1, dev.n_gpus, dev.seed, dev.device, 0, C_NULL, 0, 1, C_NULL, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ds = DeviceState(C_NULL, keep, 0)
    GC.@preserve keep begin
        st = ccall((:mavi_create, LIB), Int32, (Ref{MaviParams}, Ref{Ptr{Cvoid}}), params, h)
        ds.h = h[]
        check(ds, st)
    end
    HANDLES[system] = ds
    check(ds, ccall((:mavi_set_time, LIB), Int32, (Ptr{Cvoid}, Int64, Float64), ds.h, system.time_info.num_steps, system.time_info.time))
    if kind == 4 && !isnothing(system.info.p_neigh)      # RingsSystem(p_neighbors_cfg=...): before the upload, whose
        enable_particle_neighbors!(ds, system.info.p_neigh)   # forces! fills the lists like the reference's constructor
    end
    if kind == 4                                           # sources / sinks / VarRingsIds and invasions: before the upload too
        enable_sources!(ds, system, keep)
        enable_invasions!(ds, system)
    end
    upload_state!(system)
    finalizer(_ -> ccall((:mavi_destroy, LIB), Int32, (Ptr{Cvoid},), ds.h), ds)
    return ds
end

handle(system) = get(() -> attach!(system), HANDLES, system)

"Host state -> device (after the host edited `system.state`)."
function upload_state!(system::System)
    ds = handle(system)
    st = system.state
    mask = st.part_ids isa States.ParticleIds ? UInt8.(st.part_ids.mask) : UInt8[]
    GC.@preserve st mask check(ds, ccall((:mavi_upload_state, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int64), ds.h, pointer(st.pos), pointer(second_array(st)),
        isempty(mask) ? C_NULL : pointer(mask), length(st.pos)))
end

"`sync_to_host!(system)`: device state -> `system.state`, forces -> `get_forces(system)`.  Called by GUI / experiment / checkpoint hooks (SURVEY.md A.2)."
function sync_to_host!(system::System)
    ds = handle(system)
    st = system.state
    f = get_forces(system)
    GC.@preserve st f begin
        check(ds, ccall((:mavi_download_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ds.h, pointer(st.pos), pointer(second_array(st))))
        check(ds, ccall((:mavi_download_forces, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ds.h, pointer(f)))
    end
    return system
end

function device_step!(system::System, nsteps::Integer=1; noise=nothing)
    ds = handle(system)
    GC.@preserve noise check(ds, ccall((:mavi_step, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), ds.h, nsteps,
        isnothing(noise) ? C_NULL : pointer(noise)))
    for _ in 1:nsteps                      # update_time!, src/integration.jl:500-503 (same Float64 accumulation)
        system.time_info.time += system.int_cfg.dt
        system.time_info.num_steps += 1
    end
    ds.calls += 1
    ds.calls % system.int_cfg.device.sync_every == 0 && sync_to_host!(system)
    return nothing
end

# Every System whose IntCfg carries a CUDADevice (RingsIntCfg is the same IntCfg with `extra`, src/rings/configs.jl:344-352).
# The leading eight type parameters of `System` are spelled out WITH the bounds its definition declares
# (src/systems.jl:45-48); the trailing six (ChunksT, SysT, SpaceDataT, InfoT, DebugT, RNGT) stay free with their own bounds.
const CUDAIntCfg = IntCfg{<:Number,<:Union{ChunksCfg,Nothing},CUDADevice,<:Any}
const CUDASystem = System{T,ND,NT,StateT,WallTypeT,GeometryCfgT,DynamicCfgT,IntCfgT} where {
    T,ND,NT,StateT<:State{ND,T},WallTypeT<:WallType,GeometryCfgT<:GeometryCfg,DynamicCfgT<:DynamicCfg,IntCfgT<:CUDAIntCfg}

# ---- the dispatch seam ------------------------------------------------------------------------------------------------
newton_step!(system::CUDASystem) = device_step!(system, 1)
szabo_step!(system::CUDASystem) = device_step!(system, 1)
rtp_step!(system::CUDASystem) = device_step!(system, 1)

function calc_forces!(system::System, chunks, ::CUDADevice)
    ds = handle(system)
    check(ds, ccall((:mavi_calc_forces, LIB), Int32, (Ptr{Cvoid},), ds.h))
    f = get_forces(system)
    GC.@preserve f check(ds, ccall((:mavi_download_forces, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ds.h, pointer(f)))
end

# Mavi.Rings: `Rings.Integration.step!` (src/rings/integration.jl:522-543) and the RingsInfo fields host code reads
# (src/rings/rings.jl:118-128; SURVEY.md A.2).  `noise` = one randn per ring and step in :host_noise mode.
function Mavi.Rings.Integration.step!(system::CUDASystem; noise=nothing)
    device_step!(system, 1; noise=noise)
    if system.int_cfg.device.sync_every == 1
        sync_rings_info!(system)
        sync_rings_ids!(system)
        sync_invasions!(system)
    end
    return nothing
end

function sync_rings_info!(system)
    ds = handle(system)
    info = system.info
    GC.@preserve info check(ds, ccall((:mavi_rings_download_info, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ds.h, pointer(info.areas), pointer(info.cms), pointer(info.continuos_pos)))
    isnothing(info.p_neigh) || sync_particle_neighbors!(system)
    return info
end

"`run_system` with the default step function: ONE `mavi_step(h, nsteps)` and one download."
function run_system(system::CUDASystem; tf=nothing, num_steps=nothing, step_func=nothing)
    if !isnothing(step_func)
        return invoke(run_system, Tuple{Any}, system; tf=tf, num_steps=num_steps, step_func=step_func)
    end
    n = num_steps
    if !isnothing(tf)
        t, n = system.time_info.time, 0
        while t < tf          # the reference's own loop condition on the Float64 accumulation time += dt
            t += system.int_cfg.dt
            n += 1
        end
    end
    ds = handle(system)
    check(ds, ccall((:mavi_step, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), ds.h, n, C_NULL))
    for _ in 1:n
        system.time_info.time += system.int_cfg.dt
        system.time_info.num_steps += 1
    end
    sync_to_host!(system)
end

"(kinetic_energy, potential_energy) as device block reductions (src/quantities.jl)."
function device_energies(system::System; stencil_only=false)
    ds = handle(system)
    ke, pe = Ref(0.0), Ref(0.0)
    check(ds, ccall((:mavi_energies, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}, Ref{Float64}), ds.h, stencil_only ? 1 : 0, ke, pe))
    return ke[], pe[]
end

"""
Particle contact lists of a RingsSystem built with `p_neighbors_cfg` (src/rings/rings.jl:143-158): call
`enable_particle_neighbors!` right after `mavi_create` (inside `attach!`, before the upload) and
`sync_particle_neighbors!` wherever host code reads `system.info.p_neigh` (src/rings/neighbors.jl:58-62).
Ids cross the ABI 0-based; lists come back ascending (the reference appends in pair order and compares sorted lists).
"""
function enable_particle_neighbors!(ds::DeviceState, p_neigh)
    neigh = p_neigh.neighbors
    mode = isnothing(neigh.list) ? 1 : 2                       # MAVI_NEIGH_COUNT / MAVI_NEIGH_LIST
    check(ds, ccall((:mavi_rings_set_neighbors, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64),
                    ds.h, mode, p_neigh.type == :all ? 1 : 0, neigh.cfg.tol))
end

function sync_particle_neighbors!(system)
    ds = handle(system)
    neigh = system.info.p_neigh.neighbors
    n = size(neigh.count, 1)
    count = Vector{Int32}(undef, n)
    list = isnothing(neigh.list) ? Int32[] : Matrix{Int32}(undef, 15, n)   # num_max_neighbors = 15, column per particle
    GC.@preserve count list check(ds, ccall((:mavi_rings_download_neighbors, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}),
        ds.h, pointer(count), isempty(list) ? C_NULL : pointer(list)))
    neigh.count[:, 1] .= count
    isnothing(neigh.list) || (neigh.list[:, :, 1] .= list .+ 1)            # back to 1-based ids (padding -1 -> 0)
    return neigh
end

"""
Sources, sinks and a variable number of rings (src/rings/sources.jl, src/rings/states.jl:173-227): `RingsSystem(source_cfg=[...])`
on a `RingsState(active_state=...)`.  The list goes to the device once (`mavi_rings_set_sources`, after `mavi_create`, before the
upload); every `step!` then removes / spawns rings on the device side of the handle exactly where the reference does
(src/rings/integration.jl:353-358,523-526).  `spawn_pol = :random` draws come from Philox on the device (`spawn_draws = NULL`);
pass pre-drawn `rand(rng)` values instead for parity runs.  `sync_rings_ids!` brings `state.rings_ids` (mask, uids, ids,
num_active) back where host code reads them.
"""
struct MaviSourceSink
    kind::Int32; num_spawn_pos::Int32; spawn_pos::Ptr{Float64}
    bottom_left::NTuple{2,Float64}; spawn_pol::Float64; pad::Float64; offset::NTuple{2,Float64}; size::NTuple{2,Int32}
    sink_geom::Int32; _pad::Int32
    sink_rect_bl::NTuple{2,Float64}; sink_rect_len::Float64; sink_rect_h::Float64
    sink_circ_center::NTuple{2,Float64}; sink_circ_radius::Float64
end

function lower_source(src, keep)
    cfg = src.cfg
    if cfg isa Mavi.Rings.Sources.SourceCfg
        sp = Float64[c for q in cfg.spawn_pos for c in q]
        push!(keep, sp)
        pol = cfg.spawn_pol isa Mavi.Rings.Sources.RandomPol ? NaN : Float64(cfg.spawn_pol)
        return MaviSourceSink(0, length(cfg.spawn_pos), pointer(sp), Tuple(Float64.(cfg.bottom_left)), pol, cfg.pad,
                              Float64.(cfg.offset), Int32.(cfg.size), 0, 0, (0.0, 0.0), 0.0, 0.0, (0.0, 0.0), 0.0)
    end
    g = cfg.geometry_cfg                                   # SinkCfg
    if g isa RectangleCfg
        return MaviSourceSink(1, 0, C_NULL, (0.0, 0.0), 0.0, 0.0, (0.0, 0.0), (Int32(0), Int32(0)), 0, 0,
                              Tuple(Float64.(g.bottom_left)), g.length, g.height, (0.0, 0.0), 0.0)
    elseif g isa CircleCfg
        return MaviSourceSink(1, 0, C_NULL, (0.0, 0.0), 0.0, 0.0, (0.0, 0.0), (Int32(0), Int32(0)), 1, 0,
                              (0.0, 0.0), 0.0, 0.0, Tuple(Float64.(g.center)), g.radius)
    end
    error("SinkCfg geometry $(typeof(g)) is not supported on the device")
end

function enable_sources!(ds::DeviceState, system, keep; spawn_draws=nothing)
    ids = system.state.rings_ids
    var = ids isa Mavi.Rings.States.VarRingsIds
    srcs = system.info.sources
    (var || !isnothing(srcs)) || return
    list = isnothing(srcs) ? MaviSourceSink[] : [lower_source(s, keep) for s in srcs]
    mask = var ? UInt8.(ids.mask) : UInt8[]
    push!(keep, list); push!(keep, mask)
    draws = isnothing(spawn_draws) ? Float64[] : Float64.(spawn_draws)
    GC.@preserve list mask draws check(ds, ccall((:mavi_rings_set_sources, LIB), Int32,
        (Ptr{Cvoid}, Ptr{MaviSourceSink}, Int32, Ptr{UInt8}, Ptr{Float64}, Int64), ds.h,
        isempty(list) ? C_NULL : pointer(list), length(list), isempty(mask) ? C_NULL : pointer(mask),
        isempty(draws) ? C_NULL : pointer(draws), length(draws)))
end

function sync_rings_ids!(system)
    ids = system.state.rings_ids
    ids isa Mavi.Rings.States.VarRingsIds || return ids
    ds = handle(system)
    nr = length(ids.mask)
    mask, uids, na = Vector{UInt8}(undef, nr), Vector{Int64}(undef, nr), Ref{Int64}(0)
    GC.@preserve mask uids check(ds, ccall((:mavi_rings_download_active, LIB), Int32,
        (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}, Ref{Int64}), ds.h, pointer(mask), pointer(uids), na))
    ids.mask .= mask .!= 0
    ids.uids .= uids
    Mavi.States.update_ids!(system.state)                 # calc_active_ids! on the host copy (ids, p_ids, counts)
    return ids
end

"""
Ring invasions (src/rings/integration.jl:379-520): `RingsIntCfg(invasions_cfg=InvasionsCfg(steps_to_update), r_chunks_cfg=...)`.
The device checks every `steps_to_update` steps before the forces of that step; `sync_invasions!` refills
`system.info.invasions.list` (1-based ids, sorted; the reference lists them in pair-enumeration order).
"""
function enable_invasions!(ds::DeviceState, system)
    extra = system.int_cfg.extra
    (isnothing(extra) || isnothing(extra.invasions_cfg)) && return
    rc = extra.r_chunks_cfg
    check(ds, ccall((:mavi_rings_set_invasions, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), ds.h,
        extra.invasions_cfg.steps_to_update, isnothing(rc) ? 0 : rc.num_cols, isnothing(rc) ? 0 : rc.num_rows))
end

function sync_invasions!(system)
    inv = system.info.invasions
    extra = system.int_cfg.extra
    (isnothing(extra) || isnothing(extra.invasions_cfg)) && return inv     # invasions are off
    ds = handle(system)
    cap = 4096
    while true
        n, tri = Ref{Int64}(0), Matrix{Int32}(undef, 3, cap)
        GC.@preserve tri check(ds, ccall((:mavi_rings_download_invasions, LIB), Int32,
            (Ptr{Cvoid}, Ref{Int64}, Ptr{Int32}, Int64), ds.h, n, pointer(tri), cap))
        if n[] <= cap
            empty!(inv.list)
            for k in 1:n[]
                push!(inv.list, Mavi.Rings.Invasion(invasor=tri[1, k] + 1, invaded=tri[2, k] + 1, p_id=tri[3, k] + 1))
            end
            return inv
        end
        cap = n[]
    end
end

end # module
