/*
 * sphb200.h — C-ABI of libsphb200.so: the B200-native SPH inner loop that drops in
 * behind SPHExample's SimulationLoop / ComputeInteractions! / SimulationConstants API.
 *
 * The reference (AhmedSalih3d/SPHExample @ 54cbca9) is pure Julia and has NO FFI of its
 * own, so every entry point below names the Julia function it replaces (file:line relative
 * to the reference tree).  The Julia side binds these with `ccall` (INTEGRATION.md,
 * julia/SPHExampleB200.jl); tests, bench and the Python host mirror bind them with ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes only, no C++/torch types; every function returns 0 on success
 *    or a negative SPHB200_E* code and never throws across the boundary;
 *    sphb200_last_error() gives the message.
 *  - host arrays use exactly Julia's memory layout: Vector{SVector{D,T}} is packed T[N][D]
 *    (2D = the (x,z) columns, src/PreProcess.jl:30-34), Vector{T} is T[N]; T is float when
 *    params.real_bytes == 4 and double when 8.  The library copies; the caller keeps ownership.
 *  - one handle is used by one host thread at a time; a handle owns one CUDA device, one
 *    stream and (multi-GPU) one NCCL communicator.
 */
#ifndef SPHB200_H
#define SPHB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_ABI_VERSION 1
#define SPHB200_MAX_MOTIONS 16

/* error codes */
#define SPHB200_OK 0
#define SPHB200_EINVAL (-1)      /* bad argument / unsupported parameter combination        */
#define SPHB200_ECUDA (-2)       /* CUDA runtime error (message has the cudaError string)    */
#define SPHB200_ESTATE (-3)      /* call order violated (e.g. step before upload)            */
#define SPHB200_ECAPACITY (-4)   /* cell grid or particle capacity exceeded                  */
#define SPHB200_ENCCL (-5)       /* NCCL error / NCCL library not loadable                   */
#define SPHB200_ENUMERIC (-6)    /* non-finite time step (simulation blew up)                */

/* ParticleType enum values, src/SimulationGeometry.jl:10-14 */
#define SPHB200_FLUID 1
#define SPHB200_FIXED 2
#define SPHB200_MOVING 3

/* model selectors = the reference's dispatch singletons */
enum { SPHB200_KERNEL_WENDLANDC2 = 0, SPHB200_KERNEL_CUBICSPLINE = 1 };          /* src/SPHKernels.jl:13-18 */
enum { SPHB200_VISC_ZERO = 0, SPHB200_VISC_ARTIFICIAL = 1, SPHB200_VISC_LAMINAR = 2,
       SPHB200_VISC_LAMINAR_SPS = 3 };                                             /* src/SPHViscosityModels.jl:16-40 */
enum { SPHB200_DDT_ZERO = 0, SPHB200_DDT_ZERO_GRAVITY_LINEAR = 1, SPHB200_DDT_LINEAR = 2,
       SPHB200_DDT_COMPLEX = 3 };                                                  /* src/SPHDensityDiffusionModels.jl:31,55,107,150 */
enum { SPHB200_NO_SHIFTING = 0, SPHB200_PLANAR_SHIFTING = 1 };                     /* src/SimulationMetaDataConfiguration.jl:12-14 */
enum { SPHB200_NO_KERNEL_OUTPUT = 0, SPHB200_STORE_KERNEL_OUTPUT = 1 };            /* :16-18 */
enum { SPHB200_NO_MDBC = 0, SPHB200_SIMPLE_MDBC = 1 };                             /* :20-22 */

/* MotionDetails{D,T}, src/SimulationGeometry.jl:17-22, keyed by Geometry.GroupMarker */
typedef struct sphb200_motion {
    int64_t group_marker;
    double velocity;
    double start_time;
    double duration;
    double direction[3];
} sphb200_motion;

/* SimulationConstants{T} (src/SimulationConstantsConfiguration.jl:36-52) +
 * SPHKernelInstance{K,D,T} (src/SPHKernels.jl:30-40) + the type parameters of
 * SimulationMetaData{D,T,S,K,B,L} (src/SimulationMetaDataConfiguration.jl:28-33).
 * All scalars travel as double; the device narrows them when real_bytes == 4. */
typedef struct sphb200_params {
    int32_t abi_version;   /* SPHB200_ABI_VERSION */
    int32_t dim;           /* Dimensions: 2 or 3 */
    int32_t real_bytes;    /* FloatType: 4 or 8 */
    int32_t kernel;
    int32_t viscosity;
    int32_t diffusion;
    int32_t shifting;
    int32_t kernel_output;
    int32_t mdbc;
    int32_t n_motions;
    /* SimulationConstants */
    double rho0, dx, m0, alpha, g, c0, gamma, gamma_inv, delta_phi, cfl, cb, cb_inv, nu0,
        blin_constant, smagorinsky_constant;
    /* SPHKernelInstance */
    double k, h, h_inv, H, H_inv, H2, alphaD, eta2, cubic_eps;
    sphb200_motion motions[SPHB200_MAX_MOTIONS];
} sphb200_params;

/* What UpdateMetaData! (src/SPHCellList.jl:679-685) and the logger read after one
 * SimulationLoop call. */
typedef struct sphb200_report {
    int64_t iteration;      /* SimMetaData.Iteration (cumulative)                 */
    int64_t index_counter;  /* SimMetaData.IndexCounter = occupied cells + 1      */
    int64_t n_rebuilds;     /* UpdateNeighbors! calls so far (cumulative)          */
    int64_t n_particles;    /* particles owned by this handle                       */
    int64_t n_halo;         /* halo particles held from slab neighbours (multi-GPU) */
    double total_time;      /* SimMetaData.TotalTime                                */
    double current_dt;      /* SimMetaData.CurrentTimeStep                          */
    double delta_x;         /* Δx accumulator of the rebuild trigger                */
} sphb200_report;

typedef struct sphb200_sim sphb200_sim;

/* ---- lifetime ---------------------------------------------------------------------- */
int sphb200_abi_version(void);
/* replaces: the allocation block of RunSimulation, src/SPHCellList.jl:825-844 */
int sphb200_create(const sphb200_params *params, int device, sphb200_sim **out);
int sphb200_destroy(sphb200_sim *sim);
/* message for the last failure on this handle (sim may be NULL: last create() failure) */
const char *sphb200_last_error(const sphb200_sim *sim);

/* ---- state transfer ---------------------------------------------------------------- */
/* replaces: handing the SimParticles StructArray to RunSimulation (src/PreProcess.jl:102-116).
 * acceleration, group_marker, id, ghost_points, ghost_normals may be NULL (zeros / 1 / 1..N /
 * "no ghost node").  GravityFactor and MotionLimiter are derived from `type` exactly as
 * src/PreProcess.jl:78-98 does.  Pressure is initialised as RunSimulation does (:835). */
int sphb200_upload(sphb200_sim *sim, int64_t n, const void *position, const void *velocity,
                   const void *acceleration, const void *density, const uint8_t *type,
                   const uint64_t *group_marker, const int64_t *id, const void *ghost_points,
                   const void *ghost_normals);
/* replaces: reading SimParticles back after SimulationLoop (src/ProduceHDFVTK.jl:544-586,
 * src/SPHCellList.jl:892).  Any output pointer may be NULL.  order 0 = device cell order
 * (what the reference leaves in SimParticles after sort!), 1 = ascending ID. */
int sphb200_download(sphb200_sim *sim, int order, void *position, void *velocity,
                     void *acceleration, void *density, void *pressure, int64_t *id,
                     uint8_t *type, uint64_t *group_marker, int64_t *cells);
int64_t sphb200_num_particles(const sphb200_sim *sim);
/* overwrite the simulation clock (SimMetaData.TotalTime / Iteration), e.g. for restarts */
int sphb200_set_time(sphb200_sim *sim, double total_time, int64_t iteration);

/* ---- the hot loop ------------------------------------------------------------------ */
/* replaces: SimulationLoop(...), src/SPHCellList.jl:727-805 — one call advances
 * `while TotalTime <= next_output_time`, with Δx reset to 1+h on entry (forced rebuild, :739). */
int sphb200_simulation_loop(sphb200_sim *sim, double next_output_time, sphb200_report *report);
/* the same loop body for a fixed number of steps (parity tests, benchmarks);
 * reset_delta_x != 0 re-arms the forced rebuild as a SimulationLoop entry would. */
int sphb200_step(sphb200_sim *sim, int64_t n_steps, int reset_delta_x, sphb200_report *report);
int sphb200_get_report(sphb200_sim *sim, sphb200_report *report);
/* number of this library's kernel launches since create (bench.py's gpu_launches) */
int64_t sphb200_launch_count(const sphb200_sim *sim);
/* run all of this handle's work on the caller's CUDA stream (a cudaStream_t; NULL = the legacy
 * default stream) instead of the handle's own, so that the caller's CUDA events bracket it */
int sphb200_set_stream(sphb200_sim *sim, void *cuda_stream);
/* tuning knobs: "compact" (0/1 two-phase walk of the cull kernel), "tma" (0/1 cp.async.bulk staging),
 * "smem_kb" (shared memory per CTA of the cull kernel), "batch" (steps per host sync),
 * "generic" (1 = force the run-time-dispatched pair body), "lists" (0/1 per-particle neighbour
 * lists reused across passes), "skin" (list skin as a fraction of H), "lcap" (list entries per
 * particle), "list_smem_kb" (shared memory per CTA of the list kernel), "list_reorder" (0/1 bank-aware
 * order of the list entries), "list_local" (0/1 per-brick instead of global list maintenance; default by
 * particle count), "graph" (0/1 replay captured CUDA graphs instead of plain launches), "lean" (0/1 steps
 * that will not rebuild replay a sequence without the UpdateNeighbors! chain), "split" (-1/0/1 list kernel,
 * 2D fp64: 4 lanes per target; default by particle count), "brick_targets" (particles per brick, <= 128;
 * 0 = by particle count) */
int sphb200_set_option(sphb200_sim *sim, const char *name, double value);
/* run-time counters: "list_builds", "list_off" (1 = lists switched off after an overflow),
 * "halo_bytes_per_step", "migrated", "n_total" (owned + halo particles held), "lean_steps" / "lean_pauses"
 * (steps replayed with the lean sequence / of which had to be finished with the full one) */
int sphb200_get_stat(sphb200_sim *sim, const char *name, double *value);
/* Device time (ms, CUDA events) of the stages of ONE extra step — the reference's TimerOutputs
 * report (SimMetaData.HourGlass, labels "01" .. "12", src/SPHCellList.jl:748-800) with the labels
 * grouped the way the fused kernels group them.  ms_out[k], k < min(n, SPHB200_N_STAGES); advances
 * the simulation by one step.  In slab mode the call is collective. */
enum {
    SPHB200_STAGE_TIMESTEP = 0,   /* S0/S1 reductions (when not fused into the previous pass 2) + "01 Update TimeStep" + step control */
    SPHB200_STAGE_REBUILD = 1,    /* "02 Calculate IndexCounter" / "02a": UpdateNeighbors! (predicated on the rebuild flag) */
    SPHB200_STAGE_MOTION1 = 2,    /* "Motion" (first ProgressMotion) + the state-n snapshots pass 2 reads (Q2) */
    SPHB200_STAGE_MDBC = 3,       /* "04 Apply MDBC before Half TimeStep" */
    SPHB200_STAGE_LISTS = 4,      /* neighbour-list build + reorder (predicated; no reference stage) */
    SPHB200_STAGE_PASS1 = 5,      /* "05 First NeighborLoop" + "03 Pressure" + "06 Update To Half TimeStep" + "07 Half LimitDensityAtBoundary" */
    SPHB200_STAGE_MOTION2 = 6,    /* "Motion" (second ProgressMotion) */
    SPHB200_STAGE_PASS2 = 7,      /* "08 Second NeighborLoop" + "03" + "09 Final LimitDensityAtBoundary" + "10 Final Density" + "11 Update To Final TimeStep" (+ the next step's S0/S1 reductions) */
    SPHB200_STAGE_METADATA = 8,   /* "12 Update MetaData" */
    SPHB200_STAGE_ALLREDUCE = 9,  /* slab mode: the per-step all-reduce, including the wait for the slowest rank */
    SPHB200_STAGE_HALO1 = 10,     /* slab mode: duration of the half-step halo exchange of pass 1 (on the exchange stream) */
    SPHB200_STAGE_HALO1_EXPOSED = 11, /* ... of which not hidden behind the interior bricks */
    SPHB200_STAGE_HALO2 = 12,     /* slab mode: the same for pass 2 */
    SPHB200_STAGE_HALO2_EXPOSED = 13,
    SPHB200_N_STAGES = 16
};
int sphb200_stage_times(sphb200_sim *sim, double *ms_out, int n);

/* ---- stage-level entry points (the reference's exported step functions) ------------- */
/* UpdateNeighbors!(Particles, H⁻¹, …), src/SPHCellList.jl:138-163: hash → stable sort by cell
 * → cell ranges.  Writes IndexCounter (= occupied cells + 1) if index_counter != NULL. */
int sphb200_update_neighbors(sphb200_sim *sim, int64_t *index_counter);
/* cell table after UpdateNeighbors!: for the occupied cells in ascending (column-major) order,
 * cell coordinates cells[n_cells][D] and half-open particle ranges start[n_cells+1]
 * (UniqueCells / ParticleRanges, src/SPHCellList.jl:144-160).  Pointers may be NULL to query
 * n_cells only. */
int sphb200_get_cell_list(sphb200_sim *sim, int64_t *n_cells, int64_t *cells, int64_t *start);
/* Pressure!(P, ρ), src/SimulationEquations.jl:18-24, on state n (half=0) or n+½ (half=1) */
int sphb200_pressure(sphb200_sim *sim, int half);
/* ResetStep! + NeighborLoop! + ReductionStep!, src/SPHCellList.jl:168-217,268-317,416-484.
 * pass 0 reads (x, ρ, P, v) of state n; pass 1 reads (xₙ⁺, ρₙ⁺, P, vₙ⁺) and, like the
 * reference (Q2), ρₙ for the diffusion/viscosity terms.  Results stay on the device as
 * dρdtI / Acceleration and are also copied to drhodt_out[N] / acc_out[N][D] when non-NULL
 * (device cell order). */
int sphb200_neighbor_loop(sphb200_sim *sim, int pass, void *drhodt_out, void *acc_out);
/* Δt(Position, Velocity, Acceleration, …), src/TimeStepping.jl:24-46 */
int sphb200_delta_t(sphb200_sim *sim, double *dt);
/* ProgressMotion, src/SPHCellList.jl:575-596 */
int sphb200_progress_motion(sphb200_sim *sim, double dt2);
/* ApplyMDBCBeforeHalf!, src/SPHCellList.jl:491-505 */
int sphb200_apply_mdbc(sphb200_sim *sim);
/* HalfTimeStep + LimitDensityAtBoundary!(ρₙ⁺), src/SPHCellList.jl:624-638,781 */
int sphb200_half_time_step(sphb200_sim *sim, double dt2);
/* LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep, src/SPHCellList.jl:794-798 */
int sphb200_full_time_step(sphb200_sim *sim, double dt);
/* half-step arrays xₙ⁺ / vₙ⁺ / ρₙ⁺ (device cell order); pointers may be NULL */
int sphb200_download_half(sphb200_sim *sim, void *position_half, void *velocity_half,
                          void *density_half, void *pressure_half);
/* PlanarShifting / StoreKernelOutput side arrays (∇Cᵢ, ∇·rᵢ, ΣW, Σ∇W), device cell order */
int sphb200_download_aux(sphb200_sim *sim, void *grad_c, void *div_r, void *kernel_sum,
                         void *kernel_grad_sum);

/* ---- multi-GPU slabs (no reference counterpart: the reference is single-process) ---- */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host (torch.distributed) */
int sphb200_comm_unique_id(uint8_t id_out[128]);
/* join the slab communicator; axis = decomposition axis (0..D-1) */
int sphb200_comm_init(sphb200_sim *sim, const uint8_t id[128], int rank, int world_size, int axis);
/* slab bounds of this rank along `axis`: owns cell coordinates lo <= c < hi (INT64_MIN/MAX = open) */
int sphb200_set_slab(sphb200_sim *sim, int64_t cell_lo, int64_t cell_hi);
/* SimpleMDBC across slabs (ApplyMDBCBeforeHalf!, src/SPHCellList.jl:219-266,491-505,598-622): the
 * GhostPoints column as a global, static table — every rank passes the SAME table, ghost_points[ng][D]
 * (nonzero entries of SimParticles.GhostPoints) and the IDs of their particles, strictly ascending.
 * A ghost node is solved by the rank that owns the node's cell and the solves are all-reduced, so
 * it does not matter how far across a slab face the node lies from its particle.  Call after
 * comm_init; replaces the ghost_points argument of upload, which slab mode ignores. */
int sphb200_set_ghost_nodes(sphb200_sim *sim, int64_t ng, const void *ghost_points,
                            const int64_t *particle_ids);
/* per-cell-column particle histogram along the slab axis (for balanced slab edges):
 * counts[c - *cell_min] for c in [*cell_min, *cell_min + *n_columns) */
int sphb200_column_histogram(sphb200_sim *sim, int axis, int64_t *cell_min, int64_t *n_columns,
                             int64_t *counts, int64_t counts_capacity);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
