// tests/physics_shim.cpp — compiles the PRODUCT's per-pair formulas (sphexample_b200/csrc/
// sph_physics.cuh) as plain host C++ so that CPU tests can check them against the oracle-side
// brute force without a GPU.  Test infrastructure; not part of libsphb200.so.
#include "../include/sphb200.h"
#include "../sphexample_b200/csrc/sph_physics.cuh"

using namespace sph;

template <class T, int D>
static void run(const sphb200_params *prm, int n, const double *pos, const double *vel, const double *rho,
                const double *press, const double *rho_n, const double *vel_n, const double *ml,
                const unsigned char *a_is_i, int generic, double *drho, double *acc, double *aux) {
    Phys<T> ph = phys_from_params<T>(*prm);
    for (int a = 0; a < n; ++a) {
        PairAccum<T, D> s;
        accum_zero(s);
        T fd = T(0), fa[D];
        for (int k = 0; k < D; ++k) fa[k] = T(0);
        FastSums<T, D> fs;
        fast_zero(fs);
        FastTarget<T> ft = make_fast_target<T>(ph, (T)rho[a], (T)press[a], (T)rho_n[a], (T)ml[a], false);
        PairSide<T, D> A;
        for (int k = 0; k < D; ++k) { A.x[k] = (T)pos[a * D + k]; A.v[k] = (T)vel[a * D + k]; A.vn[k] = (T)vel_n[a * D + k]; }
        A.rho = (T)rho[a]; A.P = (T)press[a]; A.rho_n = (T)rho_n[a]; A.ml = (T)ml[a];
        for (int b = 0; b < n; ++b) {
            if (b == a) continue;
            PairSide<T, D> B;
            for (int k = 0; k < D; ++k) { B.x[k] = (T)pos[b * D + k]; B.v[k] = (T)vel[b * D + k]; B.vn[k] = (T)vel_n[b * D + k]; }
            B.rho = (T)rho[b]; B.P = (T)press[b]; B.rho_n = (T)rho_n[b]; B.ml = (T)ml[b];
            T xab[D], r2 = T(0);
            for (int k = 0; k < D; ++k) { xab[k] = A.x[k] - B.x[k]; r2 += xab[k] * xab[k]; }
            if (!(r2 <= ph.H2)) continue;
            bool role = a_is_i[(size_t)a * n + b] != 0;
            if (generic) pair_generic<T, D>(ph, A, B, xab, r2, role, s);
            else pair_fast<T, D, false>(ph, ft, xab, r2, A.v, B.v, B.rho, B.P, B.rho_n, B.ml > T(0), role, fs);
        }
        if (!generic) fast_finish<T, D>(ft, fs, fd, fa);
        drho[a] = generic ? (double)s.drho : (double)fd;
        for (int k = 0; k < D; ++k) acc[a * D + k] = generic ? (double)s.acc[k] : (double)fa[k];
        if (aux) {   // [divr, ksum, gradC[D], kgrad[D]]
            double *q = aux + (size_t)a * (2 + 2 * D);
            q[0] = (double)s.divr; q[1] = (double)s.ksum;
            for (int k = 0; k < D; ++k) { q[2 + k] = (double)s.gradC[k]; q[2 + D + k] = (double)s.kgrad[k]; }
        }
    }
}

extern "C" int shim_pair_sums(const sphb200_params *prm, int n, const double *pos, const double *vel, const double *rho,
                              const double *press, const double *rho_n, const double *vel_n, const double *ml,
                              const unsigned char *a_is_i, int generic, int use_float, double *drho, double *acc, double *aux) {
    if (prm->dim == 2 && !use_float) run<double, 2>(prm, n, pos, vel, rho, press, rho_n, vel_n, ml, a_is_i, generic, drho, acc, aux);
    else if (prm->dim == 2) run<float, 2>(prm, n, pos, vel, rho, press, rho_n, vel_n, ml, a_is_i, generic, drho, acc, aux);
    else if (!use_float) run<double, 3>(prm, n, pos, vel, rho, press, rho_n, vel_n, ml, a_is_i, generic, drho, acc, aux);
    else run<float, 3>(prm, n, pos, vel, rho, press, rho_n, vel_n, ml, a_is_i, generic, drho, acc, aux);
    return 0;
}

// per-particle updates and the EOS, for the integrator tests
extern "C" void shim_half_full(const sphb200_params *prm, int n, double *pos, double *vel, double *acc, double *rho,
                               const double *drho, const double *gf, const double *ml, double dt, double *pos_h,
                               double *vel_h, double *rho_h, int do_full) {
    Phys<double> ph = phys_from_params<double>(*prm);
    const int D = prm->dim;
    for (int i = 0; i < n; ++i) {
        if (D == 2) {
            if (!do_full) half_step<double, 2>(ph, pos + 2 * i, vel + 2 * i, acc + 2 * i, rho[i], drho[i], gf[i], ml[i], dt, pos_h + 2 * i, vel_h + 2 * i, rho_h[i]);
            else { double z[2] = {0, 0}; full_step<double, 2>(ph, pos + 2 * i, vel + 2 * i, acc + 2 * i, rho[i], drho[i], rho_h[i], gf[i], ml[i], dt, z, 0.0); }
        } else {
            if (!do_full) half_step<double, 3>(ph, pos + 3 * i, vel + 3 * i, acc + 3 * i, rho[i], drho[i], gf[i], ml[i], dt, pos_h + 3 * i, vel_h + 3 * i, rho_h[i]);
            else { double z[3] = {0, 0, 0}; full_step<double, 3>(ph, pos + 3 * i, vel + 3 * i, acc + 3 * i, rho[i], drho[i], rho_h[i], gf[i], ml[i], dt, z, 0.0); }
        }
    }
}
extern "C" double shim_eos(const sphb200_params *prm, double rho) {
    Phys<double> ph = phys_from_params<double>(*prm);
    return eos_gamma7(ph, rho);
}

// ---- brick builder (sphexample_b200/csrc/sph_bricks.h): the product's row walker on host arrays ----
#include "../sphexample_b200/csrc/sph_bricks.h"
// cell_start: dense grid of nx * nrows cells (+1), rows of nx cells; nm rows per slab layer.
// Writes at most cap bricks of row r as (t0, t1) pairs; returns the number of bricks of the row.
extern "C" int shim_row_bricks(const int *cell_start, int nx, int nm, int dim, int r, int bt, int wlimit, int *out, int cap) {
    int roff[9];
    const int nr = dim == 3 ? 9 : 3;
    const int rowbase = r * nx;
    for (int q = 0; q < nr; ++q) {
        int dm = dim == 3 ? (q % 3 - 1) : 0;
        int ds = dim == 3 ? (q / 3 - 1) : (q - 1);
        roff[q] = rowbase + (ds * nm + dm) * nx;
    }
    int n = 0;
    auto emit = [&](int t0, int t1) {
        if (n < cap) { out[2 * n] = t0; out[2 * n + 1] = t1; }
        ++n;
    };
    if (cell_start[rowbase + nx] <= cell_start[rowbase]) return 0;
    if (dim == 3) walk_row_bricks<9>(cell_start, rowbase, nx, roff, bt, wlimit, emit);
    else walk_row_bricks<3>(cell_start, rowbase, nx, roff, bt, wlimit, emit);
    return n;
}

// the masked (branch-free) fast pair body must equal the guarded one: all pairs, no cut-off test outside
extern "C" int shim_pair_sums_masked(const sphb200_params *prm, int n, const double *pos, const double *vel, const double *rho,
                                     const double *press, const double *rho_n, const double *ml, const unsigned char *a_is_i,
                                     int use_float, double *drho, double *acc) {
    auto run3 = [&](auto tag) {
        using T = decltype(tag);
        constexpr int D = 3;
        Phys<T> ph = phys_from_params<T>(*prm);
        for (int a = 0; a < n; ++a) {
            T fd = T(0), fa[D] = {T(0), T(0), T(0)};
            FastSums<T, D> fs;
            fast_zero(fs);
            T xa[D], va[D];
            for (int k = 0; k < D; ++k) { xa[k] = (T)pos[a * D + k]; va[k] = (T)vel[a * D + k]; }
            FastTarget<T> ft = make_fast_target<T>(ph, (T)rho[a], (T)press[a], (T)rho_n[a], (T)ml[a], false);
            for (int b = 0; b < n; ++b) {
                T xab[D], vb[D], r2 = T(0);
                for (int k = 0; k < D; ++k) { xab[k] = xa[k] - (T)pos[b * D + k]; vb[k] = (T)vel[b * D + k]; r2 += xab[k] * xab[k]; }
                // b == a included on purpose: the self pair contributes exact zeros on the fast path
                pair_fast<T, D, false>(ph, ft, xab, r2, va, vb, (T)rho[b], (T)press[b], (T)rho_n[b], ml[b] > 0.0,
                                       a_is_i[(size_t)a * n + b] != 0, fs);
            }
            fast_finish<T, D>(ft, fs, fd, fa);
            drho[a] = (double)fd;
            for (int k = 0; k < D; ++k) acc[a * D + k] = (double)fa[k];
        }
    };
    if (prm->dim != 3) return -1;
    if (use_float) run3(float(0)); else run3(double(0));
    return 0;
}

// ---- step control (sphexample_b200/csrc/sph_control.h): the product's one-thread logic on the host ----
#include "../sphexample_b200/csrc/sph_control.h"
static unsigned long long bits_of(double v) { unsigned long long b; memcpy(&b, &v, 8); return b; }
// Runs n steps of head (step_control) + body bookkeeping (rebuild clears the request, step_end).  Inputs per
// step: the reductions the device would have produced (max half-step displacement², visc, |a|², |v|²) and
// whether the list build of that step overflows.  trace[step] = {dt, do_rebuild, list_build, mode0, mode1,
// paused, list_move, delta_x, total_time, list_off}.
extern "C" void shim_control_trace(int n, const double *disp2, const double *visc, const double *acc2, const double *vel2,
                                   const unsigned char *build_fails, double h, double c0, double cfl, double skin,
                                   double motion_vmax, int pause, int use_float, double delta_x0, double *trace) {
    Ctl ctl;
    GridInfo grid;
    memset(&ctl, 0, sizeof ctl);
    memset(&grid, 0, sizeof grid);
    ctl.delta_x = delta_x0;
    for (int s = 0; s < n; ++s) {
        ctl.red_disp2 = bits_of(disp2[s]);
        ctl.red_visc = bits_of(visc[s]);
        ctl.red_acc2 = bits_of(acc2[s]);
        ctl.red_vel2 = bits_of(vel2[s]);
        if (use_float) step_control<float>(&ctl, &grid, (float)h, (float)c0, (float)cfl, skin, motion_vmax, pause);
        else step_control<double>(&ctl, &grid, h, c0, cfl, skin, motion_vmax, pause);
        double *t = trace + (size_t)s * 10;
        t[0] = ctl.dt; t[1] = ctl.do_rebuild; t[2] = ctl.list_build; t[3] = ctl.list_mode[0]; t[4] = ctl.list_mode[1];
        t[5] = ctl.done; t[6] = ctl.list_move; t[7] = ctl.delta_x; t[9] = ctl.list_off;
        if (ctl.done && ctl.paused) { ctl.done = 0; ctl.paused = 0; }   // the host resumes a paused step (rebuild / full sequence)
        if (ctl.do_rebuild) ctl.do_rebuild = 0;                // k_finish_rebuild
        if (ctl.list_build && build_fails[s]) ctl.list_fail = 2;
        step_end(&ctl);
        t[8] = ctl.total_time;
    }
}

// ---- bank-aware list ordering (sphexample_b200/csrc/sph_listorder.h) ----
#include "../sphexample_b200/csrc/sph_listorder.h"
// reorders one particle's list (n_slots entries, a multiple of 8, padded with indices >= total8) for lane
// phase q exactly as k_list_reorder does; returns 1 on success, 0 when the overflow scratch was too small
extern "C" int shim_rainbow_order(const unsigned short *entries, int n_slots, int q, unsigned total8, int ovf_cap,
                                  unsigned short *out) {
    unsigned short ovf[256];
    if (ovf_cap > 256) ovf_cap = 256;
    const bool ok = rainbow_order(
        n_slots, q, total8, [&](int k) -> unsigned { return entries[k]; },
        [&](int c, int u) -> unsigned short & { return out[c * 8 + u]; }, [&](int p) -> unsigned short & { return ovf[p]; }, ovf_cap);
    return ok ? 1 : 0;
}

// per-brick list validity (brick_list_decision): returns 1 when the brick must be rebuilt; *move is updated
extern "C" int shim_brick_decision(float *move, float D, double dt_prev, double dt2, double skin) {
    // (as the kernels use it: a brick that must be rebuilt now has its bound reset by the build)
    const int r = brick_list_decision(move, D, dt_prev, dt2, skin, 0.0);
    if (r == 2) *move = 0.f;
    return r == 2 ? 1 : 0;
}

// step-by-step variant of shim_control_trace for tests whose particle motion depends on the dt decided here
struct ShimCtl { Ctl ctl; GridInfo grid; };
extern "C" void *shim_ctl_new(double delta_x0) {
    ShimCtl *s = new ShimCtl;
    memset(s, 0, sizeof *s);
    s->ctl.delta_x = delta_x0;
    return s;
}
extern "C" void shim_ctl_free(void *p) { delete (ShimCtl *)p; }
extern "C" void shim_ctl_head(void *p, double disp2, double visc, double acc2, double vel2, double h, double c0, double cfl,
                              double skin, double motion_vmax, int pause, double *out) {
    ShimCtl *s = (ShimCtl *)p;
    s->ctl.red_disp2 = bits_of(disp2);
    s->ctl.red_visc = bits_of(visc);
    s->ctl.red_acc2 = bits_of(acc2);
    s->ctl.red_vel2 = bits_of(vel2);
    step_control<double>(&s->ctl, &s->grid, h, c0, cfl, skin, motion_vmax, pause);
    out[0] = s->ctl.dt; out[1] = s->ctl.do_rebuild; out[2] = s->ctl.list_build; out[3] = s->ctl.list_mode[0];
    out[4] = s->ctl.list_mode[1]; out[5] = s->ctl.done; out[6] = s->ctl.list_move; out[7] = s->ctl.delta_x;
}
extern "C" long long shim_ctl_lean_ahead(void *p, double h, double skin, long long batch) {
    return lean_steps_ahead(((ShimCtl *)p)->ctl, h, skin, batch);
}
extern "C" int shim_ctl_paused(void *p) { return ((ShimCtl *)p)->ctl.paused; }
extern "C" void shim_ctl_body(void *p) {
    ShimCtl *s = (ShimCtl *)p;
    if (s->ctl.done && s->ctl.paused) { s->ctl.done = 0; s->ctl.paused = 0; }
    s->ctl.do_rebuild = 0;
    step_end(&s->ctl);
}
