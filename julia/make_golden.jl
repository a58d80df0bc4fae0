# make_golden.jl — reference-produced golden vectors for the hot path.
#
# Runs the UNMODIFIED reference (AhmedSalih3d/SPHExample; `julia --project=<reference checkout>`)
# on the three small parity cases of this repository and writes, per case, the particle table by ID
# after ONE output interval (one call of SimulationLoop, src/SPHCellList.jl:727-805, driven by the
# reference's own RunSimulation, :808-930):
#
#     tests/golden/ref_<case>.csv        ID, Position…, Velocity…, Acceleration…, Density, Pressure
#     tests/golden/ref_<case>.meta.json  Iteration, TotalTime, CurrentTimeStep, N, constants
#
# tests/test_reference_golden.py consumes these files when present (it is skipped otherwise) and
# compares BOTH the CPU oracle (oracle/sph_oracle.cpp) and the CUDA path against them.  That turns
# "parity: unpinned" (DESIGN.md §4) into a pinned oracle: nobody could run Julia where this
# repository was built, so the vectors have to be produced once on a box that has Julia >= 1.11:
#
#     cd <SPHExample checkout>                    # the input/ paths below are relative to it
#     julia --project=. --threads=1 <this repo>/julia/make_golden.jl <this repo>/tests/golden
#
# (--threads=1 makes the reference's per-thread accumulation order deterministic; the comparison
#  tolerances in the test absorb any thread count.)  The field list follows test/runtests.jl:43-48.
# The cases, constants and interval lengths mirror tests/util.py / sphexample_b200/cases.py:
#   c1_2d     input/dam_break_2d Dp0.02 (N = 6 881), dx=0.02, c₀=88.14487860902641, δᵩ=0.1, CFL=0.2, α=0.01, Wendland k=2
#   3d_small  input/dam_break_3d Dp0.02 (N ≈ 19 k), constants of example/Dambreak3d.jl with dx = 0.02
#   c5_mdbc   input/still_wedge Dp0.02 + still_wedge_mdbc ghost nodes, constants of example/StillWedgeMDBC.jl
using SPHExample
using Printf

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
mkpath(outdir)

function dump(name, SimParticles, SimMetaData, SimConstants, SimKernel, t_out)
    order = sortperm(SimParticles.ID)
    D = length(eltype(SimParticles.Position))
    open(joinpath(outdir, "ref_$(name).csv"), "w") do io
        cols = vcat(["ID"], ["Position$(k)" for k in 0:D-1], ["Velocity$(k)" for k in 0:D-1], ["Acceleration$(k)" for k in 0:D-1],
                    ["Density", "Pressure", "Type"])
        println(io, join(cols, ","))
        for i in order
            vals = Any[SimParticles.ID[i]]
            append!(vals, SimParticles.Position[i]); append!(vals, SimParticles.Velocity[i]); append!(vals, SimParticles.Acceleration[i])
            push!(vals, SimParticles.Density[i]); push!(vals, SimParticles.Pressure[i])
            print(io, vals[1])
            for v in vals[2:end]
                @printf(io, ",%.17g", v)
            end
            println(io, ",", Int(SimParticles.Type[i]))
        end
    end
    open(joinpath(outdir, "ref_$(name).meta.json"), "w") do io
        @printf(io, "{\"case\": \"%s\", \"N\": %d, \"Iteration\": %d, \"TotalTime\": %.17g, \"CurrentTimeStep\": %.17g, \"t_out\": %.17g, ",
                name, length(SimParticles), SimMetaData.Iteration, SimMetaData.TotalTime, SimMetaData.CurrentTimeStep, t_out)
        @printf(io, "\"dx\": %.17g, \"c0\": %.17g, \"h\": %.17g, \"H\": %.17g, \"m0\": %.17g, \"threads\": %d, \"reference\": \"AhmedSalih3d/SPHExample, unmodified\"}\n",
                SimConstants.dx, SimConstants.c₀, SimKernel.h, SimKernel.H, SimConstants.m₀, Threads.nthreads())
    end
    println("wrote ref_$(name): N = ", length(SimParticles), ", iterations = ", SimMetaData.Iteration, ", t = ", SimMetaData.TotalTime)
end

# one output interval of length t_out: RunSimulation calls SimulationLoop exactly once, then stops
function run_case(name, Dimensions, geometry, consts, kernel, t_out; mdbc = false, normals = nothing)
    FloatType = Float64
    SimParticles = AllocateDataStructures(geometry)
    BMode = mdbc ? SimpleMDBC : NoMDBC
    save = mktempdir()
    meta = SimulationMetaData{Dimensions,FloatType,NoShifting,NoKernelOutput,BMode,StoreLog}(
        SimulationName = "golden_$(name)", SaveLocation = save, SimulationTime = t_out, OutputTimes = t_out,
        VisualizeInParaview = false, ExportSingleVTKHDF = true, ExportGridCells = false, OpenLogFile = false)
    logger = SimulationLogger(meta.SaveLocation; to_console = true)
    RunSimulation(SimGeometry = geometry, SimMetaData = meta, SimConstants = consts, SimKernel = kernel, SimLogger = logger,
                  SimParticles = SimParticles, SimViscosity = ArtificialViscosity(), SimDensityDiffusion = LinearDensityDiffusion(),
                  ParticleNormalsPath = normals)
    dump(name, SimParticles, meta, consts, kernel, t_out)
end

let T = Float64
    # ---- C1: 2D dam break, shipped files ----
    dx = 0.02
    g = [Geometry{2,T}(CSVFile = "./input/dam_break_2d/DamBreak2d_Dp$(dx)_Bound.csv", GroupMarker = 1, Type = Fixed, Motion = nothing);
         Geometry{2,T}(CSVFile = "./input/dam_break_2d/DamBreak2d_Dp$(dx)_Fluid.csv", GroupMarker = 2, Type = Fluid, Motion = nothing)]
    run_case("c1_2d", 2, g, SimulationConstants{T}(dx = dx, c₀ = 88.14487860902641, δᵩ = 0.1, CFL = 0.2, α = 0.01),
             SPHKernelInstance{2,T}(WendlandC2(); dx = dx, k = 2.0), 0.005)        # ≈ 55 steps
    # ---- 3D dam break, shipped Dp 0.02 files ----
    g = [Geometry{3,T}(CSVFile = "./input/dam_break_3d/DamBreak3d_Dp$(dx)_Bound.csv", GroupMarker = 1, Type = Fixed, Motion = nothing);
         Geometry{3,T}(CSVFile = "./input/dam_break_3d/DamBreak3d_Dp$(dx)_Fluid.csv", GroupMarker = 2, Type = Fluid, Motion = nothing)]
    run_case("3d_small", 3, g, SimulationConstants{T}(dx = dx, c₀ = 33.14, α = 0.1, m₀ = 1000 * dx^3, CFL = 0.2),
             SPHKernelInstance{3,T}(WendlandC2(); h = 1 * sqrt(3 * dx^2)), 0.004)   # ≈ 20 steps
    # ---- C5: StillWedge with SimpleMDBC ----
    g = [Geometry{2,T}(CSVFile = "./input/still_wedge/StillWedge_Dp$(dx)_Bound.csv", GroupMarker = 1, Type = Fixed, Motion = nothing);
         Geometry{2,T}(CSVFile = "./input/still_wedge/StillWedge_Dp$(dx)_Fluid.csv", GroupMarker = 2, Type = Fluid, Motion = nothing)]
    run_case("c5_mdbc", 2, g, SimulationConstants{T}(dx = dx, c₀ = 42.48576250492629, δᵩ = 0.1, CFL = 0.5),
             SPHKernelInstance{2,T}(WendlandC2(); dx = dx), 0.01; mdbc = true,
             normals = "./input/still_wedge_mdbc/StillWedge_Dp$(dx)_GhostNodes_Correct.csv")   # ≈ 45 steps
end
