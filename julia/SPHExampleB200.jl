# SPHExampleB200.jl — ccall binding of libsphb200.so (include/sphb200.h) for SPHExample.
#
# Drop-in site: the `SimulationLoop(...)` call inside `RunSimulation`
# (src/SPHCellList.jl:883 of AhmedSalih3d/SPHExample @ 54cbca9).  Everything around it — CSV
# loading (PreProcess.jl), logger, VTKHDF writer, ParaView launcher — stays the reference's Julia.
#
# NOTE: Julia is not installed in the build image or on the GPU box of this repository, so this
# file has been written against the header but never executed here.  It mirrors, call for call,
# the Python ctypes binding (sphexample_b200/lib.py + simulation.py) that all tests run through.
module SPHExampleB200

using StaticArrays

const libsphb200 = get(ENV, "SPHB200_LIB", joinpath(@__DIR__, "..", "sphexample_b200", "lib", "libsphb200.so"))

const SPHB200_ABI_VERSION = Int32(1)
const SPHB200_MAX_MOTIONS = 16

struct SphMotion                       # sphb200_motion  <->  MotionDetails{D,T} (src/SimulationGeometry.jl:17-22)
    group_marker::Int64
    velocity::Float64
    start_time::Float64
    duration::Float64
    direction::NTuple{3,Float64}
end
SphMotion() = SphMotion(0, 0.0, 0.0, 0.0, (0.0, 0.0, 0.0))

struct SphParams                       # sphb200_params
    abi_version::Int32
    dim::Int32
    real_bytes::Int32
    kernel::Int32
    viscosity::Int32
    diffusion::Int32
    shifting::Int32
    kernel_output::Int32
    mdbc::Int32
    n_motions::Int32
    # SimulationConstants{T} (src/SimulationConstantsConfiguration.jl:36-52)
    rho0::Float64; dx::Float64; m0::Float64; alpha::Float64; g::Float64; c0::Float64
    gamma::Float64; gamma_inv::Float64; delta_phi::Float64; cfl::Float64; cb::Float64; cb_inv::Float64
    nu0::Float64; blin_constant::Float64; smagorinsky_constant::Float64
    # SPHKernelInstance{K,D,T} (src/SPHKernels.jl:30-40)
    k::Float64; h::Float64; h_inv::Float64; H::Float64; H_inv::Float64; H2::Float64
    alphaD::Float64; eta2::Float64; cubic_eps::Float64
    motions::NTuple{SPHB200_MAX_MOTIONS,SphMotion}
end

mutable struct SphReport                # sphb200_report
    iteration::Int64
    index_counter::Int64
    n_rebuilds::Int64
    n_particles::Int64
    n_halo::Int64
    total_time::Float64
    current_dt::Float64
    delta_x::Float64
    SphReport() = new(0, 0, 0, 0, 0, 0.0, 0.0, 0.0)
end

struct SphError <: Exception
    code::Cint
    msg::String
end

last_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:sphb200_last_error, libsphb200), Cstring, (Ptr{Cvoid},), h))
check(h, rc) = rc == 0 ? nothing : throw(SphError(rc, last_error(h)))

# dispatch singletons -> selector enums of sphb200.h
kernel_id(k) = occursin("Wendland", string(typeof(k))) ? 0 : 1
viscosity_id(v) = (s = string(typeof(v)); occursin("SPS", s) ? 3 : occursin("Laminar", s) ? 2 : occursin("Artificial", s) ? 1 : 0)
diffusion_id(d) = (s = string(typeof(d)); occursin("Complex", s) ? 3 : occursin("ZeroGravityLinear", s) ? 1 : occursin("Linear", s) ? 2 : 0)
mode_is(T, name) = occursin(name, string(T))
cubic_eps(k) = hasproperty(k, :eps) ? Float64(k.eps) : 1.0    # CubicSpline{T}.eps, src/SPHKernels.jl:15-18

"""
    make_params(SimMetaData, SimConstants, SimKernel, SimViscosity, SimDensityDiffusion, SimGeometry)

Flatten the reference's configuration structs into the POD block the C-ABI takes
(the Python mirror is sphexample_b200/config.py:make_params).
"""
function make_params(SimMetaData, SimConstants, SimKernel, SimViscosity, SimDensityDiffusion, SimGeometry)
    D  = typeof(SimMetaData).parameters[1]
    T  = typeof(SimMetaData).parameters[2]
    SM, KM, BM = typeof(SimMetaData).parameters[3:5]
    motions = [SphMotion() for _ in 1:SPHB200_MAX_MOTIONS]
    nm = 0
    for geom in SimGeometry
        geom.Motion === nothing && continue
        nm += 1
        m = geom.Motion
        dir = ntuple(i -> i <= D ? Float64(m.Direction[i]) : 0.0, 3)
        motions[nm] = SphMotion(Int64(geom.GroupMarker), Float64(m.Velocity), Float64(m.StartTime), Float64(m.Duration), dir)
    end
    c, kx = SimConstants, SimKernel
    SphParams(SPHB200_ABI_VERSION, Int32(D), Int32(sizeof(T)), Int32(kernel_id(kx.kernel)), Int32(viscosity_id(SimViscosity)),
              Int32(diffusion_id(SimDensityDiffusion)), Int32(mode_is(SM, "PlanarShifting")), Int32(mode_is(KM, "StoreKernelOutput")),
              Int32(mode_is(BM, "SimpleMDBC")), Int32(nm),
              c.ρ₀, c.dx, c.m₀, c.α, c.g, c.c₀, c.γ, c.γ⁻¹, c.δᵩ, c.CFL, c.Cb, c.Cb⁻¹, c.ν₀, c.BlinConstant, c.SmagorinskyConstant,
              kx.k, kx.h, kx.h⁻¹, kx.H, kx.H⁻¹, kx.H², kx.αD, kx.η², cubic_eps(kx.kernel),
              Tuple(motions))
end

mutable struct Handle
    ptr::Ptr{Cvoid}
end

function create(params::SphParams; device::Integer = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:sphb200_create, libsphb200), Cint, (Ref{SphParams}, Cint, Ref{Ptr{Cvoid}}), params, device, out)
    rc == 0 || throw(SphError(rc, last_error(Ptr{Cvoid}(C_NULL))))
    h = Handle(out[])
    finalizer(x -> (x.ptr != C_NULL && ccall((:sphb200_destroy, libsphb200), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
    return h
end

"""Hand the SimParticles StructArray to the device.  Vector{SVector{D,T}} is packed T[N][D]
in memory, which is exactly what sphb200_upload expects; the library copies."""
function upload!(h::Handle, SimParticles)
    N = length(SimParticles)
    gm = UInt64.(SimParticles.GroupMarker)
    id = Int64.(SimParticles.ID)
    ty = reinterpret(UInt8, SimParticles.Type)
    GC.@preserve SimParticles gm id ty begin
        rc = ccall((:sphb200_upload, libsphb200), Cint,
                   (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}),
                   h.ptr, N, pointer(SimParticles.Position), pointer(SimParticles.Velocity), pointer(SimParticles.Acceleration),
                   pointer(SimParticles.Density), pointer(ty), pointer(gm), pointer(id),
                   pointer(SimParticles.GhostPoints), pointer(SimParticles.GhostNormals))
        check(h.ptr, rc)
    end
end

"""Read the particle table back into SimParticles.

`order = 1` (default): ascending ID — the order `AllocateDataStructures` leaves the table in
(`sort!(SimParticles, by = p -> p.ID)`, src/PreProcess.jl:116) and therefore the order in which the
columns this call does NOT overwrite (GravityFactor, MotionLimiter, BoundaryBool, GhostPoints,
GhostNormals, Kernel, KernelGradient, ChunkID) still sit on the host: every row stays one particle.
`order = 0` returns the device's cell order (what the reference's own `sort!(Particles, by = p ->
p.Cells)`, src/SPHCellList.jl:142, would leave behind); use it only if you also permute those
columns yourself (download the IDs first and apply `sortperm`)."""
function download!(h::Handle, SimParticles; order::Integer = 1)
    N  = length(SimParticles)
    D  = length(eltype(SimParticles.Position))
    id = Vector{Int64}(undef, N); ty = Vector{UInt8}(undef, N); gm = Vector{UInt64}(undef, N)
    cells = Matrix{Int64}(undef, D, N)
    GC.@preserve SimParticles begin
        rc = ccall((:sphb200_download, libsphb200), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}),
                   h.ptr, order, pointer(SimParticles.Position), pointer(SimParticles.Velocity), pointer(SimParticles.Acceleration),
                   pointer(SimParticles.Density), pointer(SimParticles.Pressure), id, ty, gm, cells)
        check(h.ptr, rc)
    end
    SimParticles.ID .= id
    SimParticles.GroupMarker .= gm
    SimParticles.Type .= reinterpret(eltype(SimParticles.Type), ty)
    SimParticles.Cells .= [CartesianIndex(Tuple(cells[:, i])) for i in 1:N]
    return SimParticles
end

"""
    SimulationLoop!(h, SimMetaData, SimParticles, t_next; download = true)

Replacement for `SimulationLoop(...)` (src/SPHCellList.jl:727-805): advances
`while TotalTime <= t_next` on the GPU (Δx re-armed to 1+h on entry, like :739), then updates
the fields `UpdateMetaData!` owns (:679-685) and refreshes SimParticles for the output stage.
"""
function SimulationLoop!(h::Handle, SimMetaData, SimParticles, t_next::Real; download::Bool = true)
    rep = SphReport()
    rc = ccall((:sphb200_simulation_loop, libsphb200), Cint, (Ptr{Cvoid}, Cdouble, Ref{SphReport}), h.ptr, Float64(t_next), rep)
    check(h.ptr, rc)
    SimMetaData.StepsTakenForLastOutput = rep.iteration - SimMetaData.Iteration
    SimMetaData.Iteration       = rep.iteration
    SimMetaData.TotalTime       = rep.total_time
    SimMetaData.CurrentTimeStep = rep.current_dt
    download && download!(h, SimParticles)
    g = SphReport()
    check(h.ptr, ccall((:sphb200_get_report, libsphb200), Cint, (Ptr{Cvoid}, Ref{SphReport}), h.ptr, g))
    SimMetaData.IndexCounter = g.index_counter
    return rep
end

# ---- several GPUs (one process per GPU; see INTEGRATION.md) --------------------------------------
"""
    comm_unique_id() -> Vector{UInt8}            (rank 0; broadcast the 128 bytes to the other ranks)
    comm_init!(h, id, rank, world, axis)         (axis 0 = x, 1 = y, 2 = z), then set_slab!(h, lo, hi), then upload!
"""
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:sphb200_comm_unique_id, libsphb200), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("sphb200_comm_unique_id failed ($rc): is libnccl loadable?")
    return id
end
comm_init!(h::Handle, id::Vector{UInt8}, rank::Integer, world::Integer, axis::Integer) =
    check(h.ptr, ccall((:sphb200_comm_init, libsphb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cint), h.ptr, id, rank, world, axis))
set_slab!(h::Handle, lo::Integer, hi::Integer) =
    check(h.ptr, ccall((:sphb200_set_slab, libsphb200), Cint, (Ptr{Cvoid}, Int64, Int64), h.ptr, lo, hi))

"""
    set_ghost_nodes!(h, SimParticles)

SimpleMDBC in slab mode: every rank passes the WHOLE table's nonzero `GhostPoints` rows with the IDs of
their particles, ascending by ID (call with the full, ID-sorted SimParticles, before `upload!` of the
rank's share).
"""
function set_ghost_nodes!(h::Handle, SimParticles)
    sel = findall(!iszero, SimParticles.GhostPoints)
    sel = sel[sortperm(SimParticles.ID[sel])]
    pts = SimParticles.GhostPoints[sel]
    ids = Vector{Int64}(SimParticles.ID[sel])
    GC.@preserve pts ids begin
        rc = ccall((:sphb200_set_ghost_nodes, libsphb200), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Int64}),
                   h.ptr, length(ids), pointer(pts), pointer(ids))
    end
    check(h.ptr, rc)
end

end # module
