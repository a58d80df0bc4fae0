// sph_control.h — the device-resident control block of the step loop and the one-thread logic
// that advances it.  Plain C++ (host + device): the kernels k_step_control / k_step_end
// (sph_step.cuh) are thin wrappers, and CPU tests drive the very same code (tests/physics_shim.cpp).
#pragma once

#include <limits.h>
#include <math.h>
#include <string.h>

#include "sph_physics.cuh"

namespace sph {

// ---------------------------------------------------------------------------------------------
// Device-resident control block: everything the step sequence decides on (Δt, Δx, rebuild,
// loop termination) lives here so that a whole batch of steps runs without a host round trip.
// ---------------------------------------------------------------------------------------------
struct Ctl {
    // SimulationMetaData fields owned by the loop (src/SPHCellList.jl:679-685)
    double total_time;
    double current_dt;
    double dt, dt2;            // of the step in flight
    double delta_x;            // rebuild accumulator (src/SPHCellList.jl:739-762)
    double target_time;        // SimulationLoop's next_output_time
    long long iteration;
    long long n_rebuilds;
    int use_target;            // 1: stop when total_time > target_time
    int done;                  // set by step_control when the while-condition fails
    int do_rebuild;            // this step runs UpdateNeighbors!
    int error;                 // sticky SPHB200_E* code; every kernel returns early when set
    int step_open;             // step_control ran, step_end has not
    int red_ready;             // the reductions below were produced by the fused pass-2 epilogue of the previous step (k_reduce_dt_dx skips)
    // reductions feeding Δt and Δx (bit patterns of non-negative reals, atomicMax-ed)
    unsigned long long red_disp2, red_visc, red_acc2;
    unsigned long long red_err;   // slab mode: max over ranks of -error (all-reduced with the three above)
    unsigned long long red_vel2;  // max |v|² (bounds the displacement the neighbour lists have to absorb)
    // work distribution of the interaction kernel
    int bnd_done[2];           // slab mode: boundary-layer bricks finished in pass 1 / pass 2 of this step
    int work_counter[8];       // [pass * 3 + part] (part 0 all / 1 boundary / 2 interior bricks); [6]: the list build; [7]: the list reorder
    // per-particle neighbour lists (sph_interact.cuh): which kernel serves each pass of this step
    int list_mode[2];          // LM_CULL / LM_USE
    int list_build;            // this step starts with a list build (k_list_build)
    int list_valid;            // lists exist for the current cell structure
    int list_fail;             // a build overflowed: bit 0 candidates per brick window, bit 1 entries per particle
    int list_fail_last;        // the reason of the most recent failed build (diagnostics)
    int list_off;              // lists are switched off until the next cell rebuild (after a failed build)
    int n_list_builds;
    double list_move;          // bound on any particle's displacement since the last list build
    double list_prev_vmax;     // max |v| at the previous step head
    // per-brick ("local") list maintenance, see brick_list_decision below
    int vbox_cur;              // which of the two per-cell velocity-box buffers holds the head-of-step velocities
    int bricks_flagged;        // bricks whose lists are rebuilt in this step (counted by k_list_build)
    int bricks_urgent;         // bricks whose displacement bound is used up NOW: only then does the step build at all
    double list_build_equiv;   // builds so far in units of "all bricks once" (what sphb200_get_stat("list_builds") reports)
    double vmax_now;           // max |v| at this step head (incl. moving bodies)
    long long list_missing;    // test hook (option verify_lists): pairs within H found missing from a list in use
    // lean step sequence (single GPU): the captured step holds no UpdateNeighbors! chain; a step that needs one
    // pauses itself (done = 1, paused = 1) and the host runs its body with the chain
    int paused;                // 1: paused at the step head (nothing of the body has run); 2: paused after a failed list build
                               //    (motion, mDBC and the list build have run; the passes have not)
    double last_disp4;         // what the last step head added to delta_x (4 x the largest half-step displacement)
};

enum { LIST_BUILD_NONE = 0, LIST_BUILD_ALL = 1, LIST_BUILD_FLAGGED = 2 };   // Ctl::list_build

struct GridInfo {
    int bb_min[3], bb_max[3];  // bounding box of occupied reference cells (inclusive)
    int cmin[3];               // cell coordinate of grid index 0 per axis (= bb_min - 1)
    int nx, nm, ns;            // dense grid extents: x fastest, then m, then s (slab axis)
    int ncell, nrows, nbricks;
    int nbricks_bnd;           // bricks [0, nbricks_bnd) lie in the first / last owned slab layer (slab mode)
    int own_row0, own_row1;    // rows [own_row0, own_row1) are owned by this rank (slab mode)
    int own_p0, own_p1;        // owned particle index range in sorted order
    int own_l1, own_l2;        // [own_p0, own_l1) = first owned slab layer, [own_l2, own_p1) = last one
    int n_total;               // particles on this rank (owned + halo)
};


// Which position component plays which part in the cell key ((c_s nm + c_m) nf + c_f): f = fastest
// (a row of cells runs along it), m = middle (3D only), s = slowest = the slab axis.  Default
// f = x, m = y, s = z — the reference's own cell order.  The reference orders cells with the LAST
// component most significant (CartesianIndex sort, src/SPHCellList.jl:142); that order decides the
// density-diffusion roles (Q1), see row_role().
struct AxisMap {
    int ax_f, ax_m, ax_s;
};

// Roles of the pairs between a target's cell and the cells of one neighbouring row (offsets dm, ds
// along the m / s axes).  *pre: decided by an axis that the reference compares BEFORE the fast axis
// (+1: every cell of the row is lower than the target's cell, the target is "i"; -1: higher);
// 0: the fast axis decides first, per candidate — a lower fast-axis cell is a lower cell — and for a
// candidate in the target's own fast-axis column *post decides (0: same cell, the index order does).
SPH_HD void row_role(const AxisMap &am, int D, int dm, int ds, int *pre, int *post) {
    *pre = *post = 0;
    // walk the components from most to least significant (D-1 .. 0)
    bool before_f = true;
    for (int k = D - 1; k >= 0; --k) {
        if (k == am.ax_f) {
            before_f = false;
            continue;
        }
        const int off = (k == am.ax_s) ? ds : ((D == 3 && k == am.ax_m) ? dm : 0);
        if (off == 0) continue;
        if (before_f) {
            if (*pre == 0) *pre = -off;
        } else {
            if (*post == 0) *post = -off;
        }
    }
}


#define SPH_ERR_EINVAL (-1)
#define SPH_ERR_ECUDA (-2)
#define SPH_ERR_ESTATE (-3)
#define SPH_ERR_ECAPACITY (-4)
#define SPH_ERR_ENCCL (-5)
#define SPH_ERR_ENUMERIC (-6)

SPH_HD double ctl_bits_to_double(unsigned long long b) {
    double d;
#if defined(__CUDA_ARCH__)
    d = __longlong_as_double((long long)b);
#else
    memcpy(&d, &b, 8);
#endif
    return d;
}

// One thread.  Finishes S0/S1, decides S2 (src/SPHCellList.jl:744-762) and the while-condition
// (:742).  Arithmetic is carried out in T like the reference's (its scalars are ::T).
// list_skin > 0 switches the per-particle neighbour lists on (sph_interact.cuh): a build pass
// lists every candidate within H + skin; the lists stay exact while no two particles can have
// approached by more than skin, i.e. while 2 x (bound on any particle's displacement since the
// build) <= skin.  The bound: a full step moves a particle by dt (vₙ + vₙ₊₁)/2, at most
// dt max(vmaxₙ, vmaxₙ₊₁); the half step of pass 2 by dt/2 · vₙ; moving bodies by their prescribed
// speed (motion_vmax).
// pause_on_rebuild (slab mode): a rebuild needs the host (exchange sizes), so the step that raises
// do_rebuild also raises `done`: this step's body and every later enqueued step run empty until the
// host has rebuilt and resumes the open step — steps can be enqueued in batches without a per-step
// host round trip and without ever running a step on stale cells.
// list_local: the validity of the lists is tracked per brick (brick_list_decision, k_brick_bounds)
// instead of globally: every step is a LIST_BUILD_FLAGGED build of the bricks that need one.
template <class T>
SPH_HD void step_control(Ctl *ctl, GridInfo *grid, T h, T c0, T cfl, double list_skin, double motion_vmax, int pause_on_rebuild,
                         int list_local = 0) {
    if (ctl->red_err && !ctl->error) ctl->error = -(int)ctl->red_err;   // slab mode: another rank failed
    ctl->red_err = 0ull;
    if (ctl->error) return;
    // consume the reductions unconditionally so that nothing stale survives a skipped step
    T disp = sph_sqrt((T)ctl_bits_to_double(ctl->red_disp2));
    T visc = (T)ctl_bits_to_double(ctl->red_visc);
    T acc2 = (T)ctl_bits_to_double(ctl->red_acc2);
    const double vmax = fmax(sqrt(ctl_bits_to_double(ctl->red_vel2)), motion_vmax);
    ctl->red_vel2 = 0ull;
    ctl->red_disp2 = 0ull;
    ctl->red_visc = 0ull;
    ctl->red_acc2 = 0ull;
    ctl->red_ready = 0;
    if (ctl->use_target && !(ctl->total_time <= ctl->target_time)) {
        ctl->done = 1;
        return;
    }
    if (ctl->done) return;
    ctl->delta_x = (double)((T)ctl->delta_x + T(4) * disp);
    ctl->last_disp4 = (double)(T(4) * disp);
    T dt1 = sph_sqrt(h / sph_sqrt(acc2));   // +inf when every acceleration is zero (first step)
    T dt2 = h / (c0 + visc);
    T dt = cfl * sph_min(dt1, dt2);
    if (!(dt > T(0)) || !(dt < T(1e30))) {
        ctl->error = SPH_ERR_ENUMERIC;
        return;
    }
    ctl->dt = (double)dt;
    ctl->dt2 = (double)(dt * T(0.5));
    if ((T)ctl->delta_x >= h) {
        ctl->do_rebuild = 1;
        ctl->delta_x = 0.0;
        for (int k = 0; k < 3; ++k) {
            grid->bb_min[k] = INT_MAX;
            grid->bb_max[k] = INT_MIN;
        }
    }
    ctl->red_disp2 = 0ull;
    ctl->red_visc = 0ull;
    ctl->red_acc2 = 0ull;
    for (int k = 0; k < 8; ++k) ctl->work_counter[k] = 0;
    ctl->bnd_done[0] = ctl->bnd_done[1] = 0;
    ctl->step_open = 1;
    // ---- which kernel serves the two passes of this step ---------------------------------------
    ctl->list_mode[0] = ctl->list_mode[1] = 0;   // LM_CULL
    ctl->list_build = 0;
    ctl->vmax_now = vmax;
    if (list_skin > 0.0) {
        if (ctl->list_fail) {          // the last build overflowed: no lists until the cells change
            ctl->list_fail_last = ctl->list_fail;
            ctl->list_fail = 0;
            ctl->list_valid = 0;
            ctl->list_off = 1;
        }
        if (ctl->do_rebuild) {
            ctl->list_valid = 0;
            ctl->list_off = 0;
        }
        const double margin = 0.49 * list_skin;
        const double half = ctl->dt2 * vmax;
        ctl->list_move += ctl->current_dt * fmax(ctl->list_prev_vmax, vmax);   // the step just completed
        ctl->list_prev_vmax = vmax;
        if (!ctl->list_off && list_local) {
            // local mode: lists exist for the current cells -> only the bricks whose own displacement bound
            // is used up are rebuilt (k_brick_bounds decides, k_list_build skips the others)
            ctl->list_build = ctl->list_valid ? LIST_BUILD_FLAGGED : LIST_BUILD_ALL;
            ctl->list_valid = 1;
            ctl->list_mode[0] = 2;                                  // LM_USE
            ctl->list_mode[1] = (half <= margin) ? 2 : 0;           // (implies every brick's own half-step test)
            ctl->list_move = 0.0;
        } else if (!ctl->list_off) {
            if (ctl->list_valid && ctl->list_move + half <= margin) {
                ctl->list_mode[0] = ctl->list_mode[1] = 2;          // LM_USE
            } else {
                ctl->list_build = 1;                                // k_list_build at xₙ, then both passes use it
                ctl->list_mode[0] = 2;
                ctl->list_mode[1] = (half <= margin) ? 2 : 0;
                ctl->list_move = 0.0;
                ctl->list_valid = 1;
                ctl->n_list_builds += 1;
                ctl->list_build_equiv += 1.0;
            }
        }
    }
    // pause bits: 1 = the captured sequence holds no UpdateNeighbors! chain; 2 = nor the cull kernels (a pass
    // that is not served by the lists needs the full sequence as well)
    if (((pause_on_rebuild & 1) && ctl->do_rebuild) ||
        ((pause_on_rebuild & 2) && (ctl->list_mode[0] != 2 || ctl->list_mode[1] != 2))) {
        ctl->done = 1;
        ctl->paused = 1;
    }
}

// UpdateMetaData!, src/SPHCellList.jl:679-685 (S19)
SPH_HD void step_end(Ctl *ctl) {
    if (ctl->error || ctl->done || !ctl->step_open) return;
    ctl->iteration += 1;
    ctl->current_dt = ctl->dt;
    ctl->total_time += ctl->dt;
    ctl->step_open = 0;
    ctl->red_ready = 1;   // the fused corrector of pass 2 has accumulated the Δt / Δx reductions of the new state
    ctl->vbox_cur ^= 1;   // k_cell_vbox of the next step head writes the other buffer
}

// Host side of the lean step sequence: how many of the next steps can be enqueued without the UpdateNeighbors!
// chain (and, with lists, without the cull kernels), judged from the host's copy of the control block after a
// synchronisation.  delta_x grows by about last_disp4 per step and triggers at h (step_control); the estimate
// keeps a 10 % + one step margin — a miss is not an error (the step pauses itself), only a wasted batch tail.
// 0 = take a full step next.  list_skin = skin in length units (0: no lists), batch = the most the caller enqueues.
SPH_HD long long lean_steps_ahead(const Ctl &c, double h, double list_skin, long long batch) {
    if (!c.red_ready || c.done || c.error) return 0;
    if (list_skin > 0.0) {
        // while the lists are off (after an overflow) or the half-step displacement is about to outgrow the skin
        // (pass 2 falls back to the cull kernel: list_mode[1] in step_control), full steps
        if (c.list_off || c.list_fail || !c.list_valid) return 0;
        if (c.dt2 * c.vmax_now > 0.9 * 0.49 * list_skin) return 0;
    }
    const double room = h - c.delta_x;
    if (!(room > 0.0)) return 0;
    const double d = c.last_disp4;
    if (!(d > 0.0)) return batch;
    const double k = 0.9 * room / d - 1.0;
    return k < 1.0 ? 0 : (long long)(k < (double)batch ? k : (double)batch);
}

// Per-brick list validity.  A pair (a, b) of a brick's window that is NOT in a's list was farther
// apart than H + skin when the list was built; it is missed only if the two have approached by more
// than skin since.  Over one step the relative displacement of a and b is
//   dt ((v_a + v_a')/2 - (v_b + v_b')/2),  at most  dt max(|v_a - v_b|, |v_a' - v_b'|)  <=  dt D,
// with D the diagonal of the bounding box of all velocities in the window at the two step heads (the
// boxes are kept per cell, k_cell_vbox) — a bound on RELATIVE velocities, so a water column that
// moves as a whole keeps its lists; the global rule of step_control (2 dt max|v|) cannot see that.
// `move`: accumulated bound through the previous step; the half step of pass 2 adds dt2 D.
// Returns 2 when the brick's lists must be rebuilt NOW, 1 when they will be within about `lookahead`
// more steps at the present rate (such bricks are rebuilt along with the urgent ones: a build is a
// latency-bound launch, so the bricks that are nearly due ride along instead of forcing a build of
// their own in one of the next steps), 0 otherwise.  *move is advanced, never reset here: the build
// kernel zeroes the bound of the bricks it actually rebuilds.
SPH_HD int brick_list_decision(float *move, float D, double dt_prev, double dt2, double skin, double lookahead = 0.0) {
    const float m = *move + (float)dt_prev * D * 1.0001f;   // (rounding of the accumulation never shortens the bound)
    *move = m;
    const double need = (double)m + dt2 * (double)D;
    if (need > 0.98 * skin) return 2;
    if (need + lookahead * 2.0 * dt2 * (double)D > 0.98 * skin) return 1;
    return 0;
}

}  // namespace sph
