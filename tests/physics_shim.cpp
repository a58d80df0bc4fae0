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
            else {
                FastTarget<T> ft = make_fast_target<T>(ph, A.rho, A.P, A.rho_n, A.ml, false);
                pair_fast<T, D, false>(ph, ft, xab, r2, A.v, B.v, B.rho, B.P, B.rho_n, B.ml > T(0), role, fd, fa);
            }
        }
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
