"""Independent O(N²) numpy evaluation of ComputeInteractions! — TEST INFRASTRUCTURE ONLY.

Used to pin the C++ restatement (sph_oracle.cpp) on small inputs: it shares no code and no
traversal with it (all-pairs masks instead of cell lists; explicit Q1 roles from the sorted
order), and follows the per-pair formulas of
  src/SPHCellList.jl:268-317, src/SPHKernels.jl:75-87,
  src/SPHDensityDiffusionModels.jl:100-136, src/SPHViscosityModels.jl:56-87.
WendlandC2 only; Artificial / Laminar / Zero viscosity; Linear / ZeroGravityLinear / Zero diffusion.
"""
import numpy as np


def cell_coords(pos, H_inv):
    """map_floor, src/SPHCellList.jl:56-61 (round half away from zero)."""
    return (np.sign(pos) * np.trunc(np.abs(pos) * H_inv + 0.5)).astype(np.int64)


def pair_sums(p, pos, rho, press, vel, ml, rho_n=None, order_cells=None):
    """Returns (drhodt[N], acc[N,D]) for particles ALREADY in cell-sorted order.

    rho/vel/pos are the pass inputs; rho_n is SimParticles.Density (state n, Q2) and defaults to rho.
    order_cells: cell coordinates per particle (defines Q1 roles); default from pos."""
    n, d = pos.shape
    rho_n = rho if rho_n is None else rho_n
    cells = cell_coords(pos, p.H_inv) if order_cells is None else order_cells
    idx = np.arange(n)
    xij = pos[:, None, :] - pos[None, :, :]            # x_a - x_b
    r2 = np.einsum("abk,abk->ab", xij, xij)
    mask = (r2 <= p.H2) & (idx[:, None] != idx[None, :])
    q = np.clip(np.sqrt(np.abs(r2)) * p.h_inv, 0.0, 2.0)
    fac = p.alphaD * 5 * (q - 2) ** 3 / (8 * p.h * p.h)
    gW = fac[:, :, None] * xij                          # ∇_a W_ab
    vij = vel[:, None, :] - vel[None, :, :]
    sym = np.einsum("abk,abk->ab", -vij, gW)
    cont = -rho[:, None] * (p.m0 / rho[None, :]) * sym
    # --- density diffusion with roles (Q1): V = m0/ρ of the role-"j" particle
    same = np.all(cells[:, None, :] == cells[None, :, :], axis=2)
    a_is_i = np.where(same, idx[:, None] < idx[None, :], idx[:, None] > idx[None, :])
    if p.diffusion == 0:
        D = np.zeros_like(r2)
    else:
        rho_H = 0.0
        if p.diffusion == 2:
            PH = p.rho0 * (-p.g) * -xij[:, :, -1]
            rho_H = PH * ((1.0 / (p.cb * p.gamma)) * p.rho0)
        psi_dot = 2 * ((rho_n[None, :] - rho_n[:, None]) - rho_H) * (-r2) * fac / (r2 + p.eta2)
        vol = np.where(a_is_i, p.m0 / rho_n[None, :], p.m0 / rho_n[:, None])
        D = p.delta_phi * p.h * p.c0 * vol * psi_dot
        if p.diffusion != 1:
            D = D * (ml[:, None] * ml[None, :])
    drhodt = np.where(mask, cont + D, 0.0).sum(1)
    # --- momentum
    pfac = (press[:, None] + press[None, :]) / (rho[:, None] * rho[None, :])
    um = (-p.m0 * pfac)[:, :, None] * gW
    if p.viscosity == 1:
        vdx = np.einsum("abk,abk->ab", vij, xij)
        mu = p.h * vdx / (r2 + p.eta2)
        rho_bar = 0.5 * (rho_n[:, None] + rho_n[None, :])
        pi = np.where(vdx < 0, -p.m0 * (-p.alpha * p.c0 * mu) / rho_bar, 0.0)
        um = um + pi[:, :, None] * gW
    elif p.viscosity == 2:
        xdg = np.einsum("abk,abk->ab", xij, gW)
        term = (4 * p.m0 * p.nu0 * xdg) / ((rho_n[:, None] + rho_n[None, :]) + (r2 + p.eta2))
        um = um + term[:, :, None] * vij
    acc = np.where(mask[:, :, None], um, 0.0).sum(1)
    return drhodt, acc
