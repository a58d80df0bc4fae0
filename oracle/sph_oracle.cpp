// oracle/sph_oracle.cpp — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// A CPU (fp64, C++17/OpenMP) restatement of the SPH inner loop of AhmedSalih3d/SPHExample
// @ 54cbca9, written to be the checker for libsphb200.so and the "port" CPU baseline of
// bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it; nothing under sphexample_b200/ may.
//
// It deliberately keeps the reference's *algorithmic structure* — half stencil, each unordered
// pair visited once with a symmetric scatter into per-thread private accumulators, a reduction
// afterwards, a full-table stable sort at every neighbour rebuild, serial Δt — so that it is
// independent of the GPU's gather formulation, and it keeps the reference's documented quirks
// (SURVEY.md Q1-Q14).  Citations are file:line in the reference tree.
//
// PINNING STATUS: the reference (Julia) cannot run in this environment and its own tests hold
// no golden vectors for the pair summation.  The oracle is pinned against (a) the two
// known-answer tests the reference does have (test/runtests.jl:6-16 Δt, :18-75 isolated
// particle), (b) an independent O(N²) numpy evaluation of the per-pair formulas
// (oracle/brute_force.py), (c) conservation invariants.  For cell build / pair summation /
// diffusion / viscosity / mDBC: "parity unpinned" by reference-produced outputs.
//
// Deviation from the reference, on purpose: Q5 (last cell's end index N+1 reads one element
// out of bounds under @inbounds, src/SPHCellList.jl:160,188) — the oracle ends the last cell at N.

#include "../include/sphb200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <int D>
struct Vec {
    double c[D];
    double &operator[](int k) { return c[k]; }
    const double &operator[](int k) const { return c[k]; }
};
template <int D>
inline Vec<D> vzero() {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = 0.0;
    return r;
}
template <int D>
inline Vec<D> operator-(const Vec<D> &a, const Vec<D> &b) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = a[k] - b[k];
    return r;
}
template <int D>
inline Vec<D> operator+(const Vec<D> &a, const Vec<D> &b) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = a[k] + b[k];
    return r;
}
template <int D>
inline Vec<D> operator-(const Vec<D> &a) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = -a[k];
    return r;
}
template <int D>
inline Vec<D> operator*(double s, const Vec<D> &a) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = s * a[k];
    return r;
}
template <int D>
inline Vec<D> operator*(const Vec<D> &a, double s) {
    Vec<D> r;
    for (int k = 0; k < D; ++k) r[k] = a[k] * s;
    return r;
}
template <int D>
inline double dot(const Vec<D> &a, const Vec<D> &b) {
    double s = a[0] * b[0];
    for (int k = 1; k < D; ++k) s += a[k] * b[k];
    return s;
}
template <int D>
inline bool is_zero(const Vec<D> &a) {
    for (int k = 0; k < D; ++k)
        if (a[k] != 0.0) return false;
    return true;
}

template <int D>
using Cell = std::array<int64_t, D>;

// CartesianIndex ordering is column-major: the LAST dimension is most significant.
template <int D>
inline bool cell_less(const Cell<D> &a, const Cell<D> &b) {
    for (int k = D - 1; k >= 0; --k) {
        if (a[k] != b[k]) return a[k] < b[k];
    }
    return false;
}
template <int D>
struct CellHash {
    size_t operator()(const Cell<D> &c) const {
        uint64_t hsh = 0x9E3779B97F4A7C15ull;
        for (int k = 0; k < D; ++k) {
            hsh ^= (uint64_t)c[k] + 0x9E3779B97F4A7C15ull + (hsh << 6) + (hsh >> 2);
        }
        return (size_t)hsh;
    }
};

// map_floor, src/SPHCellList.jl:56-61: sign(x) * trunc(muladd(|x|, H⁻¹, 0.5))
inline int64_t map_floor(double x, double inv_cutoff) {
    int64_t s = (x > 0.0) - (x < 0.0);
    return s * (int64_t)std::trunc(std::fma(std::fabs(x), inv_cutoff, 0.5));
}

struct OracleBase {
    virtual ~OracleBase() {}
    virtual void simulation_loop(double next_output_time) = 0;
    virtual void step(int64_t n, int reset_dx) = 0;
    virtual int64_t update_neighbors() = 0;
    virtual void pressure(int half) = 0;
    virtual void neighbor_loop(int pass) = 0;
    virtual double delta_t() = 0;
    virtual void progress_motion(double dt2) = 0;
    virtual void apply_mdbc() = 0;
    virtual void half_time_step(double dt2) = 0;
    virtual void full_time_step(double dt) = 0;
    virtual int get(const char *field, double *out) = 0;
    virtual void get_ids(int64_t *out) = 0;
    virtual void get_types(uint8_t *out) = 0;
    virtual void get_cells(int64_t *out) = 0;
    virtual int64_t get_cell_list(int64_t *cells, int64_t *start) = 0;
    virtual void get_report(sphb200_report *r) = 0;
    virtual void set_time(double t, int64_t it) = 0;
};

template <int D>
class Oracle final : public OracleBase {
    using V = Vec<D>;

  public:
    sphb200_params p;
    int64_t N;
    int nthreads;

    // --- the SimParticles table (src/PreProcess.jl:102-116); permuted by every rebuild
    std::vector<Cell<D>> cells;
    std::vector<V> pos, acc, vel, ghost, gnorm, kgrad;
    std::vector<double> rho, press, gf, ml, kern;
    std::vector<uint8_t> type;
    std::vector<int64_t> id;
    std::vector<uint64_t> group;
    // --- support arrays (src/PreProcess.jl:121-158); NOT permuted by the rebuild
    std::vector<double> drhodt, rho_h, divr;
    std::vector<V> vel_h, pos_h, gradC;
    // --- per-thread private accumulators (src/PreProcess.jl:198-215)
    std::vector<std::vector<double>> drhodt_t, divr_t, kern_t;
    std::vector<std::vector<V>> acc_t, gradC_t, kgrad_t;
    // --- cell list: UniqueCells / ParticleRanges / CellDict (src/SPHCellList.jl:840-843)
    std::vector<Cell<D>> ucells;
    std::vector<int64_t> range_start;  // [n_cells+1], 0-based half-open
    std::unordered_map<Cell<D>, int64_t, CellHash<D>> dict;
    std::vector<Cell<D>> half_stencil, full_stencil;
    // --- SimulationMetaData fields that the loop touches
    int64_t iteration = 0, index_counter = 0, n_rebuilds = 0;
    double total_time = 0.0, current_dt = 0.0, delta_x = 0.0;

    Oracle(const sphb200_params &pp, int64_t n, const double *x, const double *v, const double *a,
           const double *r, const uint8_t *ty, const uint64_t *grp, const int64_t *ids,
           const double *gp, const double *gn, int nthr)
        : p(pp), N(n), nthreads(std::max(1, nthr)) {
        cells.assign(N, Cell<D>{});
        pos.resize(N);
        acc.assign(N, vzero<D>());
        vel.assign(N, vzero<D>());
        ghost.assign(N, vzero<D>());
        gnorm.assign(N, vzero<D>());
        kgrad.assign(N, vzero<D>());
        rho.resize(N);
        press.assign(N, 0.0);
        gf.resize(N);
        ml.resize(N);
        kern.assign(N, 0.0);
        type.resize(N);
        id.resize(N);
        group.resize(N);
        for (int64_t i = 0; i < N; ++i) {
            for (int k = 0; k < D; ++k) {
                pos[i][k] = x[i * D + k];
                if (v) vel[i][k] = v[i * D + k];
                if (a) acc[i][k] = a[i * D + k];
                if (gp) ghost[i][k] = gp[i * D + k];
                if (gn) gnorm[i][k] = gn[i * D + k];
            }
            rho[i] = r[i];
            type[i] = ty[i];
            id[i] = ids ? ids[i] : i + 1;
            group[i] = grp ? grp[i] : 1;
            // GravityFactor / MotionLimiter rules, src/PreProcess.jl:78-98 (Q6)
            gf[i] = (ty[i] == SPHB200_FLUID) ? -1.0 : (ty[i] == SPHB200_MOVING ? 1.0 : 0.0);
            ml[i] = (ty[i] == SPHB200_FLUID) ? 1.0 : 0.0;
        }
        drhodt.assign(N, 0.0);
        rho_h.assign(N, 0.0);
        divr.assign(N, 0.0);
        vel_h.assign(N, vzero<D>());
        pos_h.assign(N, vzero<D>());
        gradC.assign(N, vzero<D>());
        drhodt_t.assign(nthreads, std::vector<double>(N, 0.0));
        acc_t.assign(nthreads, std::vector<V>(N, vzero<D>()));
        if (p.shifting) {
            divr_t.assign(nthreads, std::vector<double>(N, 0.0));
            gradC_t.assign(nthreads, std::vector<V>(N, vzero<D>()));
        }
        if (p.kernel_output) {
            kern_t.assign(nthreads, std::vector<double>(N, 0.0));
            kgrad_t.assign(nthreads, std::vector<V>(N, vzero<D>()));
        }
        build_stencils();
        // RunSimulation computes Pressure! once before the loop, src/SPHCellList.jl:835
        pressure(0);
    }

    // ConstructStencil, src/SPHCellList.jl:37-43: first ⌊3^D/2⌋ offsets of
    // CartesianIndices((-1:1)^D) in column-major order (first dimension fastest).
    void build_stencils() {
        int total = 1;
        for (int k = 0; k < D; ++k) total *= 3;
        for (int L = 0; L < total; ++L) {
            Cell<D> o;
            int rem = L;
            for (int k = 0; k < D; ++k) {
                o[k] = rem % 3 - 1;
                rem /= 3;
            }
            full_stencil.push_back(o);
            if (L < total / 2) half_stencil.push_back(o);
        }
    }

    // ---------------------------------------------------------------- cell list
    // UpdateNeighbors!, src/SPHCellList.jl:138-163
    int64_t update_neighbors() override {
        for (int64_t i = 0; i < N; ++i)  // ExtractCells!, :118-123
            for (int k = 0; k < D; ++k) cells[i][k] = map_floor(pos[i][k], p.H_inv);
        // sort!(Particles, by = p -> p.Cells): stable (Q13), permutes every field of the table
        std::vector<int64_t> perm(N);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
            return cell_less<D>(cells[a], cells[b]);
        });
        permute(cells, perm);
        permute(pos, perm);
        permute(acc, perm);
        permute(vel, perm);
        permute(ghost, perm);
        permute(gnorm, perm);
        permute(kgrad, perm);
        permute(rho, perm);
        permute(press, perm);
        permute(gf, perm);
        permute(ml, perm);
        permute(kern, perm);
        permute(type, perm);
        permute(id, perm);
        permute(group, perm);
        ucells.clear();
        range_start.clear();
        dict.clear();
        for (int64_t i = 0; i < N; ++i) {
            if (i == 0 || cells[i] != cells[i - 1]) {
                dict[cells[i]] = (int64_t)ucells.size();
                ucells.push_back(cells[i]);
                range_start.push_back(i);
            }
        }
        range_start.push_back(N);  // Q5: the reference writes N+1 here
        index_counter = (int64_t)ucells.size() + 1;  // IndexCounter counts the dummy entry too
        ++n_rebuilds;
        return index_counter;
    }
    template <class T>
    void permute(std::vector<T> &a, const std::vector<int64_t> &perm) {
        std::vector<T> tmp(a.size());
        for (int64_t i = 0; i < N; ++i) tmp[i] = a[perm[i]];
        a.swap(tmp);
    }

    // ---------------------------------------------------------------- kernels
    // ∇Wᵢⱼ WendlandC2, src/SPHKernels.jl:80-87; CubicSpline :94-110
    inline V grad_w(double q, const V &xij) const {
        if (p.kernel == SPHB200_KERNEL_WENDLANDC2) {
            double qm2 = q - 2.0;
            double factor = p.alphaD * 5.0 * (qm2 * qm2 * qm2) / (8.0 * p.h * p.h);
            return factor * xij;
        }
        double dwdq;
        if (q >= 0.0 && q <= 1.0)
            dwdq = p.alphaD * (-3.0 * q + (9.0 / 4.0) * (q * q));
        else if (q > 1.0 && q <= 2.0)
            dwdq = p.alphaD * (-3.0 / 4.0) * ((2.0 - q) * (2.0 - q));
        else
            dwdq = 0.0;
        double nrm = std::sqrt(dot(xij, xij));
        double s = dwdq * p.h_inv;
        V r;
        for (int k = 0; k < D; ++k) r[k] = s * xij[k] / (nrm + p.eta2);
        return r;
    }
    // Wᵢⱼ, src/SPHKernels.jl:75-78, :89-92
    inline double w(double q) const {
        if (p.kernel == SPHB200_KERNEL_WENDLANDC2) {
            double t = 1.0 - q / 2.0;
            double t2 = t * t;
            return p.alphaD * (t2 * t2) * (2.0 * q + 1.0);
        }
        double a = (q >= 0.0 && q <= 1.0) ? (1.0 - 1.5 * (q * q) + 0.75 * (q * q * q)) : 0.0;
        double b = (q > 1.0 && q <= 2.0) ? 0.25 * ((2.0 - q) * (2.0 - q) * (2.0 - q)) : 0.0;
        return p.alphaD * (a + b);
    }
    // tensile_correction, src/SPHKernels.jl:115-126
    inline double tensile(double Pi, double ri, double Pj, double rj, double q) const {
        if (p.kernel == SPHB200_KERNEL_WENDLANDC2) return 0.0;
        double ratio = w(q) / w(p.dx);
        double r2 = ratio * ratio;
        return p.cubic_eps * (((Pi / (ri * ri)) + (Pj / (rj * rj))) * (r2 * r2));
    }
    // Estimate7thRoot, src/SimulationEquations.jl:49-63
    static inline double estimate_7th_root(double x) {
        double ax = std::fabs(x);
        uint64_t bits;
        std::memcpy(&bits, &ax, 8);
        bits = 0x36cd000000000000ull + bits / 7;
        double t;
        std::memcpy(&t, &bits, 8);
        t = std::copysign(t, x);
        for (int it = 0; it < 2; ++it) {
            double t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
            double xot4 = x / t4;
            t = t - t * (t3 - xot4) / (4.0 * t3 + 3.0 * xot4);
        }
        return t;
    }

    // compute_density_diffusion, src/SPHDensityDiffusionModels.jl:31-188.
    // Reads SimParticles.Density (state n) in BOTH passes (Q2); Dⱼ = −Dᵢ with m₀/ρⱼ (Q1).
    inline double density_diffusion(const V &xij, const V &gW, double r2, int64_t i, int64_t j) const {
        if (p.diffusion == SPHB200_DDT_ZERO) return 0.0;  // Q11: scalar zero
        double rho_i = rho[i], rho_j = rho[j];
        double inv = 1.0 / (r2 + p.eta2);
        double rho_ji = rho_j - rho_i;
        double rho_H = 0.0;
        if (p.diffusion == SPHB200_DDT_LINEAR) {
            double lin = (1.0 / (p.cb * p.gamma)) * p.rho0;
            double PH = p.rho0 * (-p.g) * -xij[D - 1];
            rho_H = PH * lin;
        } else if (p.diffusion == SPHB200_DDT_COMPLEX) {
            double PH = p.rho0 * (-p.g) * -xij[D - 1];
            rho_H = p.rho0 * (estimate_7th_root(1.0 + (PH * p.cb_inv)) - 1.0);
        }
        V psi = (2.0 * (rho_ji - rho_H)) * (-xij) * inv;
        double Di = p.delta_phi * p.h * p.c0 * (p.m0 / rho_j) * dot(psi, gW);
        if (p.diffusion != SPHB200_DDT_ZERO_GRAVITY_LINEAR) Di = Di * (ml[i] * ml[j]);
        return Di;
    }

    // compute_viscosity, src/SPHViscosityModels.jl:51-126.  Reads SimParticles.Density and
    // (SPS) SimParticles.Velocity, i.e. state n, in both passes (Q2).
    inline V viscosity(const V &xij, const V &vij, const V &gW, double r2, int64_t i, int64_t j) const {
        if (p.viscosity == SPHB200_VISC_ZERO) return vzero<D>();
        double rho_i = rho[i], rho_j = rho[j];
        if (p.viscosity == SPHB200_VISC_ARTIFICIAL) {
            double vdx = dot(vij, xij);
            if (vdx < 0.0) {
                double rho_bar = 0.5 * (rho_i + rho_j);
                double mu = p.h * vdx / (r2 + p.eta2);
                return (-p.m0 * (-p.alpha * p.c0 * mu) / rho_bar) * gW;
            }
            return vzero<D>();
        }
        // Laminar (Q12: '+' between the brackets, literally)
        double term = (4.0 * p.m0 * p.nu0 * dot(xij, gW)) / ((rho_i + rho_j) + (r2 + p.eta2));
        V out = term * vij;
        if (p.viscosity == SPHB200_VISC_LAMINAR) return out;
        // LaminarSPS: per-pair SPS stress, 1/3 also in 2D (Q12)
        V vi = vel[i], vj = vel[j];
        double Si[D][D], Sj[D][D], tau[D][D];
        V dji = (p.m0 / rho_j) * (vj - vi);
        V dij = (p.m0 / rho_i) * (vi - vj);
        double si2 = 0.0, sj2 = 0.0, tri = 0.0, trj = 0.0;
        for (int a = 0; a < D; ++a)
            for (int b = 0; b < D; ++b) {
                Si[a][b] = dji[a] * gW[b];
                Sj[a][b] = dij[a] * -gW[b];
                si2 += Si[a][b] * Si[a][b];
                sj2 += Sj[a][b] * Sj[a][b];
                if (a == b) {
                    tri += Si[a][b];
                    trj += Sj[a][b];
                }
            }
        double nSi = std::sqrt(2.0 * si2), nSj = std::sqrt(2.0 * sj2);
        double csdx2 = (p.smagorinsky_constant * p.dx) * (p.smagorinsky_constant * p.dx);
        double nuti = csdx2 * nSi, nutj = csdx2 * nSj;
        for (int a = 0; a < D; ++a)
            for (int b = 0; b < D; ++b) {
                double I = (a == b) ? 1.0 : 0.0;
                double ti = 2.0 * nuti * rho_i * (Si[a][b] - (1.0 / 3.0) * tri * I) -
                            (2.0 / 3.0) * rho_i * p.blin_constant * (p.dx * p.dx) * (nSi * nSi) * I;
                double tj = 2.0 * nutj * rho_j * (Sj[a][b] - (1.0 / 3.0) * trj * I) -
                            (2.0 / 3.0) * rho_j * p.blin_constant * (p.dx * p.dx) * (nSj * nSj) * I;
                tau[a][b] = ti + tj;
            }
        double c = p.m0 / (rho_j * rho_i);
        for (int a = 0; a < D; ++a) {
            double s = 0.0;
            for (int b = 0; b < D; ++b) s += (c * tau[a][b]) * gW[b];
            out[a] += s;
        }
        return out;
    }

    // ComputeInteractions!, src/SPHCellList.jl:268-317
    inline void compute_interactions(const V *X, const double *R, const double *P, const V *U,
                                     int64_t i, int64_t j, int t) {
        V xij = X[i] - X[j];
        double r2 = dot(xij, xij);
        if (r2 <= p.H2) {
            double d = std::sqrt(std::fabs(r2));
            double q = std::min(std::max(d * p.h_inv, 0.0), 2.0);
            V gW = grad_w(q, xij);
            double rho_i = R[i], rho_j = R[j];
            V vij = U[i] - U[j];
            double sym = dot(-vij, gW);
            double drho_p = -rho_i * (p.m0 / rho_j) * sym;
            double drho_m = -rho_j * (p.m0 / rho_i) * sym;
            double Di = density_diffusion(xij, gW, r2, i, j);
            drhodt_t[t][i] += drho_p + Di;
            drhodt_t[t][j] += drho_m + (-Di);
            double Pi = P[i], Pj = P[j];
            double Pfac = (Pi + Pj) / (rho_i * rho_j);
            double fab = tensile(Pi, rho_i, Pj, rho_j, q);
            V um = (-p.m0 * (Pfac + fab)) * gW + viscosity(xij, vij, gW, r2, i, j);
            for (int k = 0; k < D; ++k) {
                acc_t[t][i][k] += um[k];
                acc_t[t][j][k] -= um[k];
            }
            if (p.kernel_output) {  // KernelOutput!, :106-116
                double W = w(q);
                kern_t[t][i] += W;
                kern_t[t][j] += W;
                for (int k = 0; k < D; ++k) {
                    kgrad_t[t][i][k] += gW[k];
                    kgrad_t[t][j][k] += -gW[k];
                }
            }
            if (p.shifting) {  // add_shifting_terms!, :73-88
                double mlc = ml[i] * ml[j];
                for (int k = 0; k < D; ++k) {
                    gradC_t[t][i][k] += (p.m0 / rho_i) * gW[k];
                    gradC_t[t][j][k] += (p.m0 / rho_j) * -gW[k];
                }
                divr_t[t][i] += (p.m0 / rho_j) * dot(-xij, gW) * mlc;
                divr_t[t][j] += (p.m0 / rho_i) * dot(xij, -gW) * mlc;
            }
        }
    }

    // ResetStep! + NeighborLoop! + ReductionStep!, src/SPHCellList.jl:168-217,416-484
    void neighbor_loop(int pass) override {
        const V *X = pass ? pos_h.data() : pos.data();
        const double *R = pass ? rho_h.data() : rho.data();
        const V *U = pass ? vel_h.data() : vel.data();
        const double *P = press.data();
        reset_step();
        int64_t ncell = (int64_t)ucells.size();
        int64_t base = (ncell + nthreads - 1) / nthreads;  // :175-176
        int64_t chunk = base + (base & 1);
        if (chunk < 1) chunk = 1;
        int64_t nchunks = (ncell + chunk - 1) / chunk;
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
        for (int64_t c = 0; c < nchunks; ++c) {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            int64_t c0 = c * chunk, c1 = std::min(c0 + chunk, ncell);
            for (int64_t it = c0; it < c1; ++it) {
                int64_t s = range_start[it], e = range_start[it + 1];
                for (int64_t i = s; i < e; ++i)  // (1) same cell, :191-196
                    for (int64_t j = i + 1; j < e; ++j) compute_interactions(X, R, P, U, i, j, t);
                for (const Cell<D> &S : half_stencil) {  // (2) half stencil, :199-210
                    Cell<D> nc;
                    for (int k = 0; k < D; ++k) nc[k] = ucells[it][k] + S[k];
                    auto f = dict.find(nc);
                    if (f == dict.end()) continue;
                    int64_t s2 = range_start[f->second], e2 = range_start[f->second + 1];
                    for (int64_t i = s; i < e; ++i)
                        for (int64_t j = s2; j < e2; ++j) compute_interactions(X, R, P, U, i, j, t);
                }
            }
        }
        reduction_step();
    }
    void reset_step() {  // :416-432
#pragma omp parallel for num_threads(nthreads)
        for (int64_t i = 0; i < N; ++i) {
            drhodt[i] = 0.0;
            acc[i] = vzero<D>();
            if (p.kernel_output) {
                kern[i] = 0.0;
                kgrad[i] = vzero<D>();
            }
            if (p.shifting) {
                gradC[i] = vzero<D>();
                divr[i] = 0.0;
            }
        }
#pragma omp parallel for num_threads(nthreads)
        for (int t = 0; t < nthreads; ++t) {
            std::fill(drhodt_t[t].begin(), drhodt_t[t].end(), 0.0);
            std::fill(acc_t[t].begin(), acc_t[t].end(), vzero<D>());
            if (p.kernel_output) {
                std::fill(kern_t[t].begin(), kern_t[t].end(), 0.0);
                std::fill(kgrad_t[t].begin(), kgrad_t[t].end(), vzero<D>());
            }
            if (p.shifting) {
                std::fill(divr_t[t].begin(), divr_t[t].end(), 0.0);
                std::fill(gradC_t[t].begin(), gradC_t[t].end(), vzero<D>());
            }
        }
    }
    void reduction_step() {  // reduce_sum!, :367-381: serial over copies, threaded over index
#pragma omp parallel for num_threads(nthreads)
        for (int64_t i = 0; i < N; ++i) {
            for (int t = 0; t < nthreads; ++t) {
                drhodt[i] += drhodt_t[t][i];
                for (int k = 0; k < D; ++k) acc[i][k] += acc_t[t][i][k];
                if (p.kernel_output) {
                    kern[i] += kern_t[t][i];
                    for (int k = 0; k < D; ++k) kgrad[i][k] += kgrad_t[t][i][k];
                }
                if (p.shifting) {
                    divr[i] += divr_t[t][i];
                    for (int k = 0; k < D; ++k) gradC[i][k] += gradC_t[t][i][k];
                }
            }
        }
    }

    // ---------------------------------------------------------------- per-particle updates
    // Pressure! / EquationOfStateGamma7, src/SimulationEquations.jl:9-24
    void pressure(int half) override {
        const double *R = half ? rho_h.data() : rho.data();
        double pre = (p.c0 * p.c0 * p.rho0) / 7.0;
        for (int64_t i = 0; i < N; ++i) {
            double x = R[i] / p.rho0;
            double x2 = x * x, x4 = x2 * x2;
            press[i] = pre * (x4 * x2 * x - 1.0);
        }
    }
    // Δt, src/TimeStepping.jl:24-46 (Q3: absolute positions, all particle types)
    double delta_t() override {
        double visc = 0.0;
        bool first = true;
        double dt1 = std::numeric_limits<double>::infinity();
        for (int64_t i = 0; i < N; ++i) {
            double val = std::fabs(p.h * dot(vel[i], pos[i]) / (dot(pos[i], pos[i]) + p.eta2));
            if (first || val > visc) visc = val;
            first = false;
            double cand = std::sqrt(p.h / std::sqrt(dot(acc[i], acc[i])));
            if (cand < dt1) dt1 = cand;
        }
        double dt2 = p.h / (p.c0 + visc);
        return p.cfl * std::min(dt1, dt2);
    }
    // update_delta_x!, src/SPHCellList.jl:706-724 (Q4: factor 4)
    double update_delta_x(double dx_acc) const {
        double maxd = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) {
                double d = pos_h[i][k] - pos[i][k];
                s += d * d;
            }
            double nrm = std::sqrt(s);
            if (nrm > maxd) maxd = nrm;
        }
        return dx_acc + 4.0 * maxd;
    }
    // ProgressMotion, src/SPHCellList.jl:575-596 (Q8)
    void progress_motion(double dt2) override {
        for (int64_t i = 0; i < N; ++i) {
            if (type[i] != SPHB200_MOVING) continue;
            const sphb200_motion *m = nullptr;
            for (int k = 0; k < p.n_motions; ++k)
                if ((uint64_t)p.motions[k].group_marker == group[i]) m = &p.motions[k];
            if (!m) continue;
            double should = (m->start_time <= total_time && total_time <= (m->start_time + m->duration)) ? 1.0 : 0.0;
            for (int k = 0; k < D; ++k) {
                vel[i][k] = m->velocity * m->direction[k] * should;
                pos[i][k] += vel[i][k] * dt2;
            }
        }
    }
    // HalfTimeStep + LimitDensityAtBoundary!(ρₙ⁺), src/SPHCellList.jl:624-638,781
    void half_time_step(double dt2) override {
        for (int64_t i = 0; i < N; ++i) {
            acc[i][D - 1] += p.g * gf[i];
            for (int k = 0; k < D; ++k) {
                pos_h[i][k] = pos[i][k] + vel[i][k] * dt2 * ml[i];
                vel_h[i][k] = vel[i][k] + acc[i][k] * dt2 * ml[i];
            }
            rho_h[i] = rho[i] + drhodt[i] * dt2;
        }
        limit_density(rho_h);
    }
    void limit_density(std::vector<double> &R) {  // src/SimulationEquations.jl:36-42
        for (int64_t i = 0; i < N; ++i)
            if (R[i] < p.rho0 && ml[i] == 0.0) R[i] = p.rho0;
    }
    // LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep, src/SPHCellList.jl:794-798
    void full_time_step(double dt) override {
        limit_density(rho);
        for (int64_t i = 0; i < N; ++i) {  // DensityEpsi!, src/SimulationEquations.jl:28-33
            double epsi = -(drhodt[i] / rho_h[i]) * dt;
            rho[i] *= (2.0 - epsi) / (2.0 + epsi);
        }
        for (int64_t i = 0; i < N; ++i) {  // FullTimeStep, :640-677
            acc[i][D - 1] += p.g * gf[i];
            V adt = acc[i] * dt * ml[i];
            for (int k = 0; k < D; ++k) vel[i][k] += adt[k];
            V shift = vzero<D>();
            if (p.shifting) {
                double fsc = (divr[i] - 0.0) / ((double)D - 0.0);
                if (!(fsc < 0.0)) {
                    double vn = std::sqrt(dot(vel[i], vel[i]));
                    shift = (-fsc * 2.0 * p.h * vn * dt) * gradC[i];
                }
            }
            for (int k = 0; k < D; ++k)
                pos[i][k] += (((vel[i][k] + (vel[i][k] - adt[k])) / 2.0) * dt + shift[k]) * ml[i];
        }
    }

    // ---------------------------------------------------------------- mDBC
    // ApplyMDBCBeforeHalf!, src/SPHCellList.jl:219-266,319-365,491-505,598-622
    void apply_mdbc() override {
        constexpr int E = D + 1;
        // two phases like the reference (all (b, A) first, then ApplyMDBCCorrection)
        std::vector<double> new_rho(N);
        std::vector<uint8_t> has_new(N, 0);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
        for (int64_t i = 0; i < N; ++i) {
            if (is_zero<D>(ghost[i])) continue;  // Q10 sentinel
            double b[E] = {0}, A[E][E] = {{0}};
            Cell<D> gc;
            for (int k = 0; k < D; ++k) gc[k] = map_floor(ghost[i][k], p.H_inv);
            for (const Cell<D> &S : full_stencil) {
                Cell<D> nc;
                for (int k = 0; k < D; ++k) nc[k] = gc[k] + S[k];
                auto f = dict.find(nc);
                if (f == dict.end()) continue;
                for (int64_t j = range_start[f->second]; j < range_start[f->second + 1]; ++j) {
                    if (type[j] != SPHB200_FLUID) continue;
                    V xij = ghost[i] - pos[j];
                    double r2 = dot(xij, xij);
                    if (r2 > p.H2) continue;
                    double d = std::sqrt(std::fabs(r2));
                    double q = std::min(std::max(d * p.h_inv, 0.0), 2.0);
                    double W = w(q);
                    V gW = grad_w(q, xij);
                    double Vj = p.m0 / rho[j];
                    double col[E];
                    col[0] = Vj * W;
                    b[0] += p.m0 * W;
                    for (int k = 0; k < D; ++k) {
                        col[k + 1] = Vj * gW[k];
                        b[k + 1] += p.m0 * gW[k];
                    }
                    // column-major fill: column 0 = col, column c = xⱼᵢ[c-1] * col
                    for (int r = 0; r < E; ++r) {
                        A[r][0] += col[r];
                        for (int c = 1; c < E; ++c) A[r][c] += (-xij[c - 1]) * col[r];
                    }
                }
            }
            double detA = det<E>(A);
            if (std::fabs(detA) >= 1e-3) {
                double sol[E];
                solve<E>(A, b, sol);
                double v1 = sol[0];
                for (int k = 0; k < D; ++k) v1 += sol[k + 1] * (pos[i][k] - ghost[i][k]);
                new_rho[i] = std::isnan(v1) ? p.rho0 : v1;
                has_new[i] = 1;
            } else if (A[0][0] > 0.0) {
                double v = b[0] / A[0][0];
                new_rho[i] = std::isnan(v) ? p.rho0 : v;
                has_new[i] = 1;
            }
        }
        for (int64_t i = 0; i < N; ++i)
            if (has_new[i]) rho[i] = new_rho[i];
    }
    template <int E>
    static double det(const double A[E][E]) {
        double M[E][E];
        std::memcpy(M, A, sizeof(M));
        double d = 1.0;
        for (int c = 0; c < E; ++c) {
            int piv = c;
            for (int r = c + 1; r < E; ++r)
                if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
            if (M[piv][c] == 0.0) return 0.0;
            if (piv != c) {
                for (int k = 0; k < E; ++k) std::swap(M[piv][k], M[c][k]);
                d = -d;
            }
            d *= M[c][c];
            for (int r = c + 1; r < E; ++r) {
                double f = M[r][c] / M[c][c];
                for (int k = c; k < E; ++k) M[r][k] -= f * M[c][k];
            }
        }
        return d;
    }
    template <int E>
    static void solve(const double A[E][E], const double b[E], double x[E]) {
        double M[E][E + 1];
        for (int r = 0; r < E; ++r) {
            for (int c = 0; c < E; ++c) M[r][c] = A[r][c];
            M[r][E] = b[r];
        }
        for (int c = 0; c < E; ++c) {
            int piv = c;
            for (int r = c + 1; r < E; ++r)
                if (std::fabs(M[r][c]) > std::fabs(M[piv][c])) piv = r;
            if (piv != c)
                for (int k = 0; k <= E; ++k) std::swap(M[piv][k], M[c][k]);
            for (int r = c + 1; r < E; ++r) {
                double f = M[r][c] / M[c][c];
                for (int k = c; k <= E; ++k) M[r][k] -= f * M[c][k];
            }
        }
        for (int r = E - 1; r >= 0; --r) {
            double s = M[r][E];
            for (int c = r + 1; c < E; ++c) s -= M[r][c] * x[c];
            x[r] = s / M[r][r];
        }
    }

    // ---------------------------------------------------------------- the loop
    // one iteration of the while body of SimulationLoop, src/SPHCellList.jl:742-802
    void one_step() {
        delta_x = update_delta_x(delta_x);  // S0
        double dt = delta_t();              // S1
        double dt2 = dt * 0.5;
        if (delta_x >= p.h) {  // S2
            update_neighbors();
            delta_x = 0.0;
        }
        progress_motion(dt2);           // S3
        pressure(0);                    // S5 (S4 ResetStep! is inside neighbor_loop)
        if (p.mdbc) apply_mdbc();       // S6
        neighbor_loop(0);               // S7+S8
        half_time_step(dt2);            // S9+S10
        progress_motion(dt2);           // S12
        pressure(1);                    // S13
        neighbor_loop(1);               // S14+S15
        full_time_step(dt);             // S16-S18
        iteration += 1;                 // S19, UpdateMetaData! :679-685
        current_dt = dt;
        total_time += dt;
    }
    void simulation_loop(double next_output_time) override {
        delta_x = 1.0 + p.h;  // :739
        while (total_time <= next_output_time) one_step();
    }
    void step(int64_t n, int reset_dx) override {
        if (reset_dx) delta_x = 1.0 + p.h;
        for (int64_t s = 0; s < n; ++s) one_step();
    }

    // ---------------------------------------------------------------- getters
    int get(const char *field, double *out) override {
        std::string f(field);
        auto putv = [&](const std::vector<V> &a) {
            for (int64_t i = 0; i < N; ++i)
                for (int k = 0; k < D; ++k) out[i * D + k] = a[i][k];
            return 0;
        };
        auto puts = [&](const std::vector<double> &a) {
            std::copy(a.begin(), a.end(), out);
            return 0;
        };
        if (f == "pos") return putv(pos);
        if (f == "vel") return putv(vel);
        if (f == "acc") return putv(acc);
        if (f == "pos_h") return putv(pos_h);
        if (f == "vel_h") return putv(vel_h);
        if (f == "gradC") return putv(gradC);
        if (f == "kgrad") return putv(kgrad);
        if (f == "ghost") return putv(ghost);
        if (f == "rho") return puts(rho);
        if (f == "press") return puts(press);
        if (f == "drhodt") return puts(drhodt);
        if (f == "rho_h") return puts(rho_h);
        if (f == "divr") return puts(divr);
        if (f == "kern") return puts(kern);
        return -1;
    }
    void get_ids(int64_t *out) override { std::copy(id.begin(), id.end(), out); }
    void get_types(uint8_t *out) override { std::copy(type.begin(), type.end(), out); }
    void get_cells(int64_t *out) override {
        for (int64_t i = 0; i < N; ++i)
            for (int k = 0; k < D; ++k) out[i * D + k] = cells[i][k];
    }
    int64_t get_cell_list(int64_t *cs, int64_t *start) override {
        int64_t nc = (int64_t)ucells.size();
        if (cs)
            for (int64_t c = 0; c < nc; ++c)
                for (int k = 0; k < D; ++k) cs[c * D + k] = ucells[c][k];
        if (start) std::copy(range_start.begin(), range_start.end(), start);
        return nc;
    }
    void get_report(sphb200_report *r) override {
        r->iteration = iteration;
        r->index_counter = index_counter;
        r->n_rebuilds = n_rebuilds;
        r->n_particles = N;
        r->n_halo = 0;
        r->total_time = total_time;
        r->current_dt = current_dt;
        r->delta_x = delta_x;
    }
    void set_time(double t, int64_t it) override {
        total_time = t;
        iteration = it;
    }
};

}  // namespace

extern "C" {

void *orc_create(const sphb200_params *p, int64_t n, const double *pos, const double *vel,
                 const double *acc, const double *rho, const uint8_t *type, const uint64_t *group,
                 const int64_t *id, const double *ghost, const double *gnorm, int nthreads) {
    if (!p || n < 1 || !pos || !rho || !type) return nullptr;
    if (p->dim == 2) return new Oracle<2>(*p, n, pos, vel, acc, rho, type, group, id, ghost, gnorm, nthreads);
    if (p->dim == 3) return new Oracle<3>(*p, n, pos, vel, acc, rho, type, group, id, ghost, gnorm, nthreads);
    return nullptr;
}
void orc_destroy(void *h) { delete static_cast<OracleBase *>(h); }
void orc_simulation_loop(void *h, double t) { static_cast<OracleBase *>(h)->simulation_loop(t); }
void orc_step(void *h, int64_t n, int reset_dx) { static_cast<OracleBase *>(h)->step(n, reset_dx); }
int64_t orc_update_neighbors(void *h) { return static_cast<OracleBase *>(h)->update_neighbors(); }
void orc_pressure(void *h, int half) { static_cast<OracleBase *>(h)->pressure(half); }
void orc_neighbor_loop(void *h, int pass) { static_cast<OracleBase *>(h)->neighbor_loop(pass); }
double orc_delta_t(void *h) { return static_cast<OracleBase *>(h)->delta_t(); }
void orc_progress_motion(void *h, double dt2) { static_cast<OracleBase *>(h)->progress_motion(dt2); }
void orc_apply_mdbc(void *h) { static_cast<OracleBase *>(h)->apply_mdbc(); }
void orc_half_time_step(void *h, double dt2) { static_cast<OracleBase *>(h)->half_time_step(dt2); }
void orc_full_time_step(void *h, double dt) { static_cast<OracleBase *>(h)->full_time_step(dt); }
int orc_get(void *h, const char *field, double *out) { return static_cast<OracleBase *>(h)->get(field, out); }
void orc_get_ids(void *h, int64_t *out) { static_cast<OracleBase *>(h)->get_ids(out); }
void orc_get_types(void *h, uint8_t *out) { static_cast<OracleBase *>(h)->get_types(out); }
void orc_get_cells(void *h, int64_t *out) { static_cast<OracleBase *>(h)->get_cells(out); }
int64_t orc_get_cell_list(void *h, int64_t *cells, int64_t *start) {
    return static_cast<OracleBase *>(h)->get_cell_list(cells, start);
}
void orc_get_report(void *h, sphb200_report *r) { static_cast<OracleBase *>(h)->get_report(r); }
void orc_set_time(void *h, double t, int64_t it) { static_cast<OracleBase *>(h)->set_time(t, it); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
