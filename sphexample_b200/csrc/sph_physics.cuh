// sph_physics.cuh — per-pair and per-particle SPH formulas of the hot path, written ONCE for the
// device kernels (and compilable as plain host C++ so tests can exercise them without a GPU).
//
// Gather formulation: every function evaluates the contribution of neighbour `b` to particle
// `a` (x_ab = x_a - x_b, grad = ∇_a W_ab).  The reference visits each unordered pair once and
// scatters to both ends (src/SPHCellList.jl:268-317); all of its pair terms are antisymmetric /
// symmetric under the swap except the density-diffusion volume factor, which is
// handled by the `a_is_i` role flag.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SPH_HD __host__ __device__ __forceinline__
#else
#define SPH_HD inline
#endif

namespace sph {

enum { K_WENDLAND = 0, K_CUBIC = 1 };
enum { V_ZERO = 0, V_ARTIFICIAL = 1, V_LAMINAR = 2, V_SPS = 3 };
enum { DDT_ZERO = 0, DDT_ZG_LINEAR = 1, DDT_LINEAR = 2, DDT_COMPLEX = 3 };

// SimulationConstants + SPHKernelInstance narrowed to the working precision, plus a few
// products that are constant over a run.
template <class T>
struct Phys {
    T rho0, m0, alpha, g, c0, gamma, delta_phi, cb, cb_inv, nu0, blin, smag, dx;
    T h, h_inv, H2, alphaD, eta2, cubic_eps;
    T gradw_c;   // alphaD * 5 / (8 h^2)          WendlandC2 ∇W prefactor, src/SPHKernels.jl:80-87
    T ddt_k;     // delta_phi * h * c0             src/SPHDensityDiffusionModels.jl:132
    T ddt_lin;   // rho0 * g * rho0 / (cb * gamma)  ρᴴ = ddt_lin * z_ab, :113-122
    T eos_b;     // c0^2 rho0 / 7                  src/SimulationEquations.jl:9-11
    T w_dx;      // W(dx) for the tensile correction, src/SPHKernels.jl:120-126
    T visc_k2;   // 2 m0 alpha c0 h : Π coefficient with 1/ρ̄ = 2/(ρ_a+ρ_b), src/SPHViscosityModels.jl:66-70
    int kernel, viscosity, diffusion, shifting, kernel_output, mdbc;
};

template <class T> SPH_HD T sph_sqrt(T x) { return sqrt(x); }
template <> SPH_HD float sph_sqrt<float>(float x) { return sqrtf(x); }
template <class T> SPH_HD T sph_abs(T x) { return fabs(x); }
template <> SPH_HD float sph_abs<float>(float x) { return fabsf(x); }
// reciprocal / square root of the fast pair body: IEEE on the host and in fp64, MUFU in fp32
template <class T> SPH_HD T sph_rcp(T x) { return T(1) / x; }
template <> SPH_HD float sph_rcp<float>(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
template <class T> SPH_HD T sph_sqrt_fast(T x) { return sqrt(x); }
template <> SPH_HD float sph_sqrt_fast<float>(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
template <class T> SPH_HD T sph_min(T a, T b) { return a < b ? a : b; }
template <class T> SPH_HD T sph_max(T a, T b) { return a > b ? a : b; }

// EquationOfStateGamma7, src/SimulationEquations.jl:9-11
template <class T>
SPH_HD T eos_gamma7(const Phys<T> &p, T rho) {
    T x = rho / p.rho0;
    T x2 = x * x, x4 = x2 * x2;
    return p.eos_b * (x4 * x2 * x - T(1));
}

// Wᵢⱼ, src/SPHKernels.jl:75-78 (Wendland C2), :89-92 (cubic spline)
template <class T>
SPH_HD T kernel_w(const Phys<T> &p, T q) {
    if (p.kernel == K_WENDLAND) {
        T t = T(1) - q / T(2);
        T t2 = t * t;
        return p.alphaD * (t2 * t2) * (T(2) * q + T(1));
    }
    T a = (q >= T(0) && q <= T(1)) ? (T(1) - T(1.5) * (q * q) + T(0.75) * (q * q * q)) : T(0);
    T b = (q > T(1) && q <= T(2)) ? T(0.25) * ((T(2) - q) * (T(2) - q) * (T(2) - q)) : T(0);
    return p.alphaD * (a + b);
}

// Estimate7thRoot, src/SimulationEquations.jl:49-63 (bit trick defined on Float64 only; the
// fp32 build evaluates it in double and narrows, as the trick cannot be restated on 32 bits)
SPH_HD double estimate_7th_root(double x) {
    double ax = fabs(x);
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(ax);
#else
    memcpy(&bits, &ax, 8);
#endif
    bits = 0x36cd000000000000ull + bits / 7;
    double t;
#if defined(__CUDA_ARCH__)
    t = __longlong_as_double((long long)bits);
#else
    memcpy(&t, &bits, 8);
#endif
    t = copysign(t, x);
    for (int it = 0; it < 2; ++it) {
        double t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
        double xot4 = x / t4;
        t = t - t * (t3 - xot4) / (4.0 * t3 + 3.0 * xot4);
    }
    return t;
}

// One neighbour's state as the pair loop sees it.
template <class T, int D>
struct PairSide {
    T x[D];    // position of the pass (x or xₙ⁺)
    T v[D];    // velocity of the pass (v or vₙ⁺)
    T rho;     // density of the pass (ρ or ρₙ⁺): continuity + pressure terms
    T P;       // pressure of the pass
    T rho_n;   // SimParticles.Density = state n: diffusion + viscosity terms (Q2)
    T ml;      // MotionLimiter (1 fluid, 0 boundary)
    T vn[D];   // SimParticles.Velocity = state n: LaminarSPS only (Q2)
};

template <class T, int D>
struct PairAccum {
    T drho;
    T acc[D];
    // optional modes (PlanarShifting / StoreKernelOutput), src/SPHCellList.jl:65-116
    T gradC[D];
    T divr;
    T ksum;
    T kgrad[D];
};

template <class T, int D>
SPH_HD void accum_zero(PairAccum<T, D> &s) {
    s.drho = T(0);
    s.divr = T(0);
    s.ksum = T(0);
    for (int k = 0; k < D; ++k) s.acc[k] = s.gradC[k] = s.kgrad[k] = T(0);
}

// ---------------------------------------------------------------------------------------------
// FAST path: WendlandC2 + ArtificialViscosity + LinearDensityDiffusion, no shifting / kernel
// output — the model set of every BASELINE config (example/Dambreak3d.jl:57-59 etc.).
// xab / r2 are passed in because the caller has just computed them for the cut-off test.
//
// The reference's per-pair expressions (src/SPHCellList.jl:268-317) are regrouped so that every
// factor that depends on the target particle only leaves the neighbour sum:
//   dρ/dt|a = [αD 5/(8h²)] · ( ρ_a m0 · S1  −  2 δφ h c0 m0 ML_a · S2 )
//   dv/dt|a = [αD 5/(8h²)] · (m0/ρ_a) · S3
//     S1 = Σ_b (q−2)³ (1/ρ_b) v_ab·x_ab                                     continuity, :288-291
//     S2 = Σ_b fluid_b (q−2)³ V_ab (ρₙ_b − ρₙ_a − ρᴴ_ab) r²/(r²+η²),   V_ab = 1/ρₙ of the "j" role (Q1, Q2)
//     S3 = Σ_b (q−2)³ [ (2 α c0 h ρ_a/m0)… min(v_ab·x_ab,0) / ((r²+η²)(ρₙ_a+ρₙ_b)) − (P_a+P_b)/ρ_b ] x_ab
// (∇W = αD 5 (q−2)³/(8h²) x_ab, src/SPHKernels.jl:80-87).  The two reciprocals 1/(r²+η²) and
// 1/(ρₙ_a+ρₙ_b) come from ONE reciprocal of their product.  Per pair: sqrt, 1/ρ_b, that reciprocal
// (+ 1/ρₙ_b in pass 2) — on the device in fp32 single MUFU approximations (1-2 ulp), IEEE in fp64
// and on the host; see the tolerances in tests/test_gpu_parity.py.
// ---------------------------------------------------------------------------------------------
template <class T>
struct FastTarget {
    T P, rhon, inv_rhon;     // P_a, ρₙ_a, 1/ρₙ_a
    T c_visc;                // 2 α c0 h ρ_a            (visc_k2 ρ_a / m0)
    T k_cont, k_ddt, k_acc;  // scales applied once to S1, S2, S3 (fast_finish)
};
template <class T, int D>
struct FastSums {
    T s1, s2, s3[D];
};
template <class T, int D>
SPH_HD void fast_zero(FastSums<T, D> &s) {
    s.s1 = s.s2 = T(0);
    for (int k = 0; k < D; ++k) s.s3[k] = T(0);
}
template <class T>
SPH_HD FastTarget<T> make_fast_target(const Phys<T> &p, T rho, T P, T rhon, T ml, bool same_rho) {
    FastTarget<T> f;
    f.P = P;
    f.rhon = rhon;
    T inv_rho = T(1) / rho;
    f.inv_rhon = same_rho ? inv_rho : T(1) / rhon;
    f.c_visc = (p.visc_k2 / p.m0) * rho;
    f.k_cont = p.gradw_c * (rho * p.m0);
    f.k_ddt = p.gradw_c * (T(-2) * p.ddt_k * p.m0 * ml);
    f.k_acc = p.gradw_c * (p.m0 * inv_rho);
    return f;
}
template <class T, int D>
SPH_HD void fast_finish(const FastTarget<T> &a, const FastSums<T, D> &s, T &drho, T *acc) {
    drho = a.k_cont * s.s1 + a.k_ddt * s.s2;
    for (int k = 0; k < D; ++k) acc[k] = a.k_acc * s.s3[k];
}

// SAME_RHO: the pass density is the state-n density (pass 1 of the step), so 1/ρ_b serves both
// the continuity/pressure terms and the diffusion volume.
// Branch-free: (q−2) is clamped at 0, so a candidate beyond the support (q >= 2, i.e. r² > H² up to
// the rounding of sqrt) contributes exact zeros to every sum and callers may feed candidates
// without testing r² <= H² first.  fp64 keeps the reference's exact acceptance test.
template <class T, int D, bool SAME_RHO>
SPH_HD void pair_fast(const Phys<T> &p, const FastTarget<T> &a, const T *xab, T r2, const T *va, const T *vb, T rho_b,
                      T P_b, T rhon_b, bool fluid_b, bool a_is_i, FastSums<T, D> &s) {
    T d = sph_sqrt_fast(r2);
    T qm2;
    if (sizeof(T) == 8) {
        qm2 = sph_min(d * p.h_inv, T(2)) - T(2);
        qm2 = (r2 <= p.H2) ? qm2 : T(0);
    } else {
        qm2 = sph_min(d * p.h_inv - T(2), T(0));
    }
    T fac = qm2 * (qm2 * qm2);
    T vdotx = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) vdotx += (va[k] - vb[k]) * xab[k];
    T inv_rho_b = sph_rcp(rho_b);
    T inv_rhon_b = SAME_RHO ? inv_rho_b : sph_rcp(rhon_b);
    s.s1 += (inv_rho_b * fac) * vdotx;
    T sum_rhon = a.rhon + rhon_b;
    T ip = sph_rcp((r2 + p.eta2) * sum_rhon);   // 1 / ((r²+η²)(ρₙ_a+ρₙ_b))
    T inv = ip * sum_rhon;                      // 1 / (r²+η²)
    T diff = (rhon_b - a.rhon) - p.ddt_lin * xab[D - 1];
    T dd = (a_is_i ? inv_rhon_b : a.inv_rhon) * (diff * ((fac * r2) * inv));
    s.s2 += fluid_b ? dd : T(0);
    T c1 = a.c_visc * (sph_min(vdotx, T(0)) * ip) - (a.P + P_b) * inv_rho_b;
    T cf = c1 * fac;
#pragma unroll
    for (int k = 0; k < D; ++k) s.s3[k] += cf * xab[k];
}

// ---------------------------------------------------------------------------------------------
// GENERIC path: every kernel / viscosity / diffusion / mode combination the reference dispatches
// on (src/SPHKernels.jl, src/SPHViscosityModels.jl, src/SPHDensityDiffusionModels.jl,
// src/SPHCellList.jl:65-116), selected by uniform run-time switches.
// ---------------------------------------------------------------------------------------------
template <class T, int D>
SPH_HD void pair_generic(const Phys<T> &p, const PairSide<T, D> &a, const PairSide<T, D> &b, const T *xab,
                         T r2, bool a_is_i, PairAccum<T, D> &s) {
    T d = sph_sqrt(sph_abs(r2));
    T q = sph_min(sph_max(d * p.h_inv, T(0)), T(2));
    T gW[D];
    if (p.kernel == K_WENDLAND) {   // src/SPHKernels.jl:80-87
        T qm2 = q - T(2);
        T fac = p.gradw_c * (qm2 * qm2 * qm2);
#pragma unroll
        for (int k = 0; k < D; ++k) gW[k] = fac * xab[k];
    } else {                        // CubicSpline, :94-110
        T dwdq;
        if (q >= T(0) && q <= T(1))
            dwdq = p.alphaD * (T(-3) * q + T(2.25) * (q * q));
        else if (q > T(1) && q <= T(2))
            dwdq = p.alphaD * T(-0.75) * ((T(2) - q) * (T(2) - q));
        else
            dwdq = T(0);
        T sc = dwdq * p.h_inv;
#pragma unroll
        for (int k = 0; k < D; ++k) gW[k] = sc * xab[k] / (d + p.eta2);
    }
    T vab[D];
    T vdotg = T(0), vdotx = T(0), xdotg = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        vab[k] = a.v[k] - b.v[k];
        vdotg += vab[k] * gW[k];
        vdotx += vab[k] * xab[k];
        xdotg += xab[k] * gW[k];
    }
    T inv = T(1) / (r2 + p.eta2);
    T mlab = a.ml * b.ml;
    // continuity
    T drho = a.rho * (p.m0 / b.rho) * vdotg;
    // density diffusion (Q1, Q2, Q11)
    if (p.diffusion != DDT_ZERO) {
        // Evaluated in the orientation of the reference's visit (i -> j) and negated for the
        // "j" end (Dⱼ = −Dᵢ): the Complex model's ρᴴ is not an odd function of z_ij, so the
        // two orientations are not interchangeable (src/SPHDensityDiffusionModels.jl:169-171).
        T zij = a_is_i ? xab[D - 1] : -xab[D - 1];
        T rho_ji = a_is_i ? (b.rho_n - a.rho_n) : (a.rho_n - b.rho_n);
        T rho_H = T(0);
        if (p.diffusion == DDT_LINEAR) {
            rho_H = p.ddt_lin * zij;
        } else if (p.diffusion == DDT_COMPLEX) {   // :150-188
            T PH = p.rho0 * p.g * zij;
            rho_H = p.rho0 * (T(estimate_7th_root(double(T(1) + PH * p.cb_inv))) - T(1));
        }
        T psi_dot_g = T(2) * (rho_ji - rho_H) * (-xdotg) * inv;   // (−x_ij)·∇ᵢW_ij is orientation-invariant
        T vol = p.m0 / (a_is_i ? b.rho_n : a.rho_n);
        T Dd = p.ddt_k * vol * psi_dot_g;
        if (!a_is_i) Dd = -Dd;
        if (p.diffusion != DDT_ZG_LINEAR) Dd *= mlab;
        drho += Dd;
    }
    s.drho += drho;
    // momentum
    T fab = T(0);
    if (p.kernel != K_WENDLAND) {   // tensile_correction, src/SPHKernels.jl:120-126
        T ratio = kernel_w(p, q) / p.w_dx;
        T r2_ = ratio * ratio;
        fab = p.cubic_eps * ((a.P / (a.rho * a.rho) + b.P / (b.rho * b.rho)) * (r2_ * r2_));
    }
    T coef = -p.m0 * ((a.P + b.P) / (a.rho * b.rho) + fab);
    T um[D];
#pragma unroll
    for (int k = 0; k < D; ++k) um[k] = coef * gW[k];
    if (p.viscosity == V_ARTIFICIAL) {
        if (vdotx < T(0)) {
            T rho_bar = T(0.5) * (a.rho_n + b.rho_n);
            T mu = p.h * vdotx * inv;
            T pi = p.m0 * (p.alpha * p.c0 * mu) / rho_bar;
#pragma unroll
            for (int k = 0; k < D; ++k) um[k] += pi * gW[k];
        }
    } else if (p.viscosity == V_LAMINAR || p.viscosity == V_SPS) {
        // Q12: '+' between the two brackets, literally (src/SPHViscosityModels.jl:85)
        T term = (T(4) * p.m0 * p.nu0 * xdotg) / ((a.rho_n + b.rho_n) + (r2 + p.eta2));
#pragma unroll
        for (int k = 0; k < D; ++k) um[k] += term * vab[k];
        if (p.viscosity == V_SPS) {   // per-pair SPS stress, :90-126; symmetric under a<->b
            T dba[D], dab[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                dba[k] = (p.m0 / b.rho_n) * (b.vn[k] - a.vn[k]);   // builds S of "a"
                dab[k] = (p.m0 / a.rho_n) * (a.vn[k] - b.vn[k]);   // builds S of "b" (with -∇W)
            }
            T sa2 = T(0), sb2 = T(0), tra = T(0), trb = T(0);
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    T Sa = dba[r] * gW[c], Sb = dab[r] * -gW[c];
                    sa2 += Sa * Sa;
                    sb2 += Sb * Sb;
                    if (r == c) {
                        tra += Sa;
                        trb += Sb;
                    }
                }
            T nSa = sph_sqrt(T(2) * sa2), nSb = sph_sqrt(T(2) * sb2);
            T csdx2 = (p.smag * p.dx) * (p.smag * p.dx);
            T nuta = csdx2 * nSa, nutb = csdx2 * nSb;
            T cc = p.m0 / (b.rho_n * a.rho_n);
#pragma unroll
            for (int r = 0; r < D; ++r) {
                T sres = T(0);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    T I = (r == c) ? T(1) : T(0);
                    T Sa = dba[r] * gW[c], Sb = dab[r] * -gW[c];
                    T ta = T(2) * nuta * a.rho_n * (Sa - (T(1) / T(3)) * tra * I) -
                           (T(2) / T(3)) * a.rho_n * p.blin * (p.dx * p.dx) * (nSa * nSa) * I;
                    T tb = T(2) * nutb * b.rho_n * (Sb - (T(1) / T(3)) * trb * I) -
                           (T(2) / T(3)) * b.rho_n * p.blin * (p.dx * p.dx) * (nSb * nSb) * I;
                    sres += (cc * (ta + tb)) * gW[c];
                }
                um[r] += sres;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) s.acc[k] += um[k];
    if (p.kernel_output) {   // KernelOutput!, src/SPHCellList.jl:106-116
        s.ksum += kernel_w(p, q);
#pragma unroll
        for (int k = 0; k < D; ++k) s.kgrad[k] += gW[k];
    }
    if (p.shifting) {        // add_shifting_terms!, :73-88
#pragma unroll
        for (int k = 0; k < D; ++k) s.gradC[k] += (p.m0 / a.rho) * gW[k];
        s.divr += (p.m0 / b.rho) * (-xdotg) * mlab;
    }
}

// ---------------------------------------------------------------------------------------------
// per-particle updates
// ---------------------------------------------------------------------------------------------
// HalfTimeStep + LimitDensityAtBoundary!(ρₙ⁺), src/SPHCellList.jl:624-638,781 (S9, S10)
template <class T, int D>
SPH_HD void half_step(const Phys<T> &p, const T *x, const T *v, T *acc, T rho, T drho, T gf, T ml, T dt2,
                      T *xh, T *vh, T &rhoh) {
    acc[D - 1] += p.g * gf;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        xh[k] = x[k] + v[k] * dt2 * ml;
        vh[k] = v[k] + acc[k] * dt2 * ml;
    }
    rhoh = rho + drho * dt2;
    if (rhoh < p.rho0 && ml == T(0)) rhoh = p.rho0;
}

// LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep, src/SPHCellList.jl:794-798,640-677,
// src/SimulationEquations.jl:28-42 (S16-S18).  shift_* only with PlanarShifting.
template <class T, int D>
SPH_HD void full_step(const Phys<T> &p, T *x, T *v, T *acc, T &rho, T drho, T rhoh, T gf, T ml, T dt,
                      const T *gradC, T divr) {
    if (rho < p.rho0 && ml == T(0)) rho = p.rho0;
    T epsi = -(drho / rhoh) * dt;
    rho *= (T(2) - epsi) / (T(2) + epsi);
    acc[D - 1] += p.g * gf;
    T adt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        adt[k] = acc[k] * dt * ml;
        v[k] += adt[k];
    }
    T shift[D];
#pragma unroll
    for (int k = 0; k < D; ++k) shift[k] = T(0);
    if (p.shifting) {
        T fsc = (divr - T(0)) / (T(D) - T(0));
        if (!(fsc < T(0))) {
            T vn2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) vn2 += v[k] * v[k];
            T sc = -fsc * T(2) * p.h * sph_sqrt(vn2) * dt;
#pragma unroll
            for (int k = 0; k < D; ++k) shift[k] = sc * gradC[k];
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] += (((v[k] + (v[k] - adt[k])) / T(2)) * dt + shift[k]) * ml;
}

// Narrow the C-ABI parameter block (include/sphb200.h) to the working precision.  P is
// sphb200_params; a template parameter only to keep this header free of the ABI include.
template <class T, class P>
inline Phys<T> phys_from_params(const P &p) {
    Phys<T> ph;
    ph.rho0 = (T)p.rho0; ph.m0 = (T)p.m0; ph.alpha = (T)p.alpha; ph.g = (T)p.g; ph.c0 = (T)p.c0;
    ph.gamma = (T)p.gamma; ph.delta_phi = (T)p.delta_phi; ph.cb = (T)p.cb; ph.cb_inv = (T)p.cb_inv;
    ph.nu0 = (T)p.nu0; ph.blin = (T)p.blin_constant; ph.smag = (T)p.smagorinsky_constant; ph.dx = (T)p.dx;
    ph.h = (T)p.h; ph.h_inv = (T)p.h_inv; ph.H2 = (T)p.H2; ph.alphaD = (T)p.alphaD; ph.eta2 = (T)p.eta2;
    ph.cubic_eps = (T)p.cubic_eps;
    ph.gradw_c = (T)(p.alphaD * 5.0 / (8.0 * p.h * p.h));
    ph.ddt_k = (T)(p.delta_phi * p.h * p.c0);
    ph.ddt_lin = (T)(p.rho0 * p.g * ((1.0 / (p.cb * p.gamma)) * p.rho0));
    ph.eos_b = (T)((p.c0 * p.c0 * p.rho0) / 7.0);
    ph.visc_k2 = (T)(2.0 * p.m0 * p.alpha * p.c0 * p.h);
    ph.kernel = p.kernel; ph.viscosity = p.viscosity; ph.diffusion = p.diffusion;
    ph.shifting = p.shifting; ph.kernel_output = p.kernel_output; ph.mdbc = p.mdbc;
    ph.w_dx = T(1);
    ph.w_dx = kernel_w(ph, (T)p.dx);   // Wᵢⱼ(SimKernel, dx), src/SPHKernels.jl:121
    return ph;
}

// ---------------------------------------------------------------------------------------------
// mDBC ghost-node system, src/SPHCellList.jl:319-365,598-622
// ---------------------------------------------------------------------------------------------
template <int E>
SPH_HD double det_lu(const double (&A)[E][E]) {
    double M[E][E];
    for (int r = 0; r < E; ++r)
        for (int c = 0; c < E; ++c) M[r][c] = A[r][c];
    double dt = 1.0;
    for (int c = 0; c < E; ++c) {
        int piv = c;
        for (int r = c + 1; r < E; ++r)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0) return 0.0;
        if (piv != c) {
            for (int k = 0; k < E; ++k) {
                double t = M[piv][k];
                M[piv][k] = M[c][k];
                M[c][k] = t;
            }
            dt = -dt;
        }
        dt *= M[c][c];
        for (int r = c + 1; r < E; ++r) {
            double f = M[r][c] / M[c][c];
            for (int k = c; k < E; ++k) M[r][k] -= f * M[c][k];
        }
    }
    return dt;
}

template <int E>
SPH_HD void solve_lu(const double (&A)[E][E], const double (&b)[E], double (&x)[E]) {
    double M[E][E + 1];
    for (int r = 0; r < E; ++r) {
        for (int c = 0; c < E; ++c) M[r][c] = A[r][c];
        M[r][E] = b[r];
    }
    for (int c = 0; c < E; ++c) {
        int piv = c;
        for (int r = c + 1; r < E; ++r)
            if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (piv != c)
            for (int k = 0; k <= E; ++k) {
                double t = M[piv][k];
                M[piv][k] = M[c][k];
                M[c][k] = t;
            }
        for (int r = c + 1; r < E; ++r) {
            double f = M[r][c] / M[c][c];
            for (int k = c; k <= E; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int r = E - 1; r >= 0; --r) {
        double sacc = M[r][E];
        for (int c = r + 1; c < E; ++c) sacc -= M[r][c] * x[c];
        x[r] = sacc / M[r][r];
    }
}

}  // namespace sph
