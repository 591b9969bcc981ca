"""Oracle (test infrastructure): the named residual compositions the fused kernels are
checked against, written with the restated reference operators only.

* ``advection_residual`` / ``advection_spectral`` follow ``test/advection.jl:52-83``.
* ``dissipation_residual`` / ``dissipation_spectral`` follow ``test/dissipation.jl:54-77``.
* ``euler_residual`` / ``euler_ghost_update`` are OUR canonical Euler composition
  (SURVEY.md Appendix A.10 -- the reference ships no Euler residual); they mirror the
  structure of ``test/advection.jl:67-83`` with the ``src/cfd.jl`` fluxes and ``FlowBC``.
"""
import numpy as np

from . import cfd, turbulence
from .domain import (JST_sensor, MUSCL, at_faces, cell_gradient, face_distance, face_gradient, green_gauss,
                     impose_bc, unsigned_green_gauss)

F32 = np.float32


def advection_spectral(part, C):
    """Per-cell CFL denominator of ``test/advection.jl:52-59`` (max over dims)."""
    s = None
    for dim in range(part.ndims):
        t = unsigned_green_gauss(part, at_faces(part, C[:, dim], dim), dim)
        s = t if s is None else np.maximum(s, t)
    return s


def advection_residual(part, u, ud, C):
    """Closure body of ``march!``, ``test/advection.jl:67-83``; ``C`` is (cells_in_part, nd)."""
    D = JST_sensor(part, u)
    for dim in range(part.ndims):
        Cf = at_faces(part, C[:, dim], dim)
        gu = cell_gradient(part, u, dim)
        uL, uR = MUSCL(part, u, gu, dim, D=D, high_order=True)
        ud -= green_gauss(part, (uL + uR) * Cf / 2 + np.abs(Cf) * (uL - uR) / 2, dim)


def dissipation_spectral(part):
    """Per-cell denominator of ``test/dissipation.jl:54-61``."""
    s = None
    for dim in range(part.ndims):
        t = unsigned_green_gauss(part, F32(1.0) / face_distance(part, dim), dim)
        s = t if s is None else s + t
    return s


def dissipation_residual(part, uv, uvd):
    """Closure body of ``march!``, ``test/dissipation.jl:69-77``."""
    for dim in range(part.ndims):
        uvd += green_gauss(part, face_gradient(part, uv, dim), dim)


def euler_residual(fluid, flux="hll"):
    """Canonical Euler residual closure ``f(part, Q, R, cfl)``."""

    def f(part, Q, R, cfl):
        P = cfd.state2primitive(fluid, Q)
        D = JST_sensor(part, P[:, 0])
        a = cfd.speed_of_sound(fluid, P[:, 1])
        R[...] = 0
        cfl[...] = 0
        for dim in range(part.ndims):
            gP = cell_gradient(part, P, dim)
            PL, PR = MUSCL(part, P, gP, dim, D=D, high_order=False)
            if flux == "hll":
                Fx = cfd.inviscid_fluxes_hll(fluid, PL, PR, dim)
            else:
                Fx = cfd.inviscid_fluxes_sensor(fluid, PL, PR, at_faces(part, D, dim), at_faces(part, D, dim), dim)
            R[...] = R - green_gauss(part, Fx, dim)  # float64 flux rounds into the float32 residual here
            cfl += unsigned_green_gauss(part, np.abs(at_faces(part, P[:, 2 + dim], dim)) + at_faces(part, a, dim), dim)

    return f


def rans_residual(fluid, sigma_R=F32(0.72), C1=F32(0.0829), kappa=F32(0.41)):
    """Canonical RANS residual closure ``f(part, Q, qR, R, RR, cfl)`` for configuration C5 (OURS, SURVEY.md A.10 extended:
    the reference ships no composed residual, only the pieces).  ``Q`` is the conservative mean-flow state (N, nd + 2),
    ``qR = rho * R`` the transported Wray-Agarwal variable (``src/turbulence.jl:197-218``: ``nu_t = R``).

    1. ``R, cfl`` = the Euler residual above (MUSCL + JST blend + HLL, ``green_gauss``), unchanged;
    2. for every dim: ``R += green_gauss(viscous_fluxes(fluid, at_faces(P), face_gradient(P, dim, cell_gradient(P)), dim;
       mu_t = at_faces(rho R)))`` (``src/cfd.jl:664-736``, ``src/ImmersedBoundary.jl:1039-1069``);
    3. ``RR = sum_dim green_gauss(rho_f (nu + nu_R)_f face_gradient(R) - rho_f (u_f (R_L + R_R) / 2 - |u_f| (R_R - R_L) / 2))
       + rho S`` with ``(nu_t, nu_R, S) = Wray_Agarwal(R, shear_rate(grad u), grad R, grad shear_rate)``
       (``R_t = -div(u R) + div[(nu + nu_R) grad R] + S``, ``src/turbulence.jl:213-218``), MUSCL without a sensor."""
    euler = euler_residual(fluid)

    def f(part, Q, qR, R, RR, cfl):
        nd = part.ndims
        euler(part, Q, R, cfl)
        P = cfd.state2primitive(fluid, Q)
        rho = Q[:, 0]
        Rt = qR / rho
        gP = cell_gradient(part, P)
        gR = cell_gradient(part, Rt)
        vg = [[gP[j][:, 2 + i] for j in range(nd)] for i in range(nd)]       # vg[i][j] = d u_i / d x_j
        Sr = turbulence.shear_rate(vg)
        gS = cell_gradient(part, Sr)
        wa = turbulence.wray_agarwal(Rt, Sr, np.stack(gR, axis=1), np.stack(gS, axis=1), sigma_R, C1, kappa)
        mu_t = rho * wa["nu_t"]
        nu_eff = cfd.dynamic_viscosity(fluid, P[:, 1]) / rho + wa["nu_R"]
        RR[...] = 0
        for dim in range(nd):
            Fv = cfd.viscous_fluxes(fluid, at_faces(part, P, dim), face_gradient(part, P, dim, gP), dim,
                                    mu_t=at_faces(part, mu_t, dim))
            R[...] = R + green_gauss(part, Fv, dim)
            RL, RRt = MUSCL(part, Rt, gR[dim], dim)
            uf = at_faces(part, P[:, 2 + dim], dim)
            rf = at_faces(part, rho, dim)
            Fc = rf * (uf * (RL + RRt) / F32(2) - np.abs(uf) * (RRt - RL) / F32(2))
            Fd = rf * at_faces(part, nu_eff, dim) * face_gradient(part, Rt, dim)
            RR[...] = RR + green_gauss(part, Fd - Fc, dim)
        RR[...] = RR + rho * wa["S"]

    return f


def rans_ghost_update(dom, Q, qR, R_bcs):
    """IB ghost update of the transported variable of C5: for every (boundary name, value) in order,
    ``R = qR ./ rho; impose_bc!(dom, name, R) do b, Ri; value end; qR[ghosts] = rho[ghosts] .* R[ghosts]`` (Jacobi within
    a family; ``R = 0`` at walls, ``R_inf = 3 nu`` in the far field, ``src/turbulence.jl:203``)."""
    for name, val in R_bcs:
        Rt = (qR / Q[:, 0]).astype(F32)
        impose_bc(lambda b, Ri: F32(val), dom, name, Rt)
        g = np.unique(np.concatenate([b.ghost_indices for b in dom.boundaries[name].values()] or [np.zeros(0, np.int64)]))
        qR[g] = Q[g, 0] * Rt[g]


def euler_ghost_update(dom, fluid, Q, bcs):
    """IB ghost update on the conservative state.

    ``bcs`` is an ordered list of (boundary name, FlowBC).  For every entry, in order:
    ``P = state2primitive(Q); impose_bc!(dom, name, P) do b, Pi; bc(Pi, b.normals) end;
    Q[ghosts] = primitive2state(P[ghosts])`` -- each boundary is a self-contained Q -> Q map,
    applied Jacobi-style (all image reads before any ghost write).
    """
    touched = []
    for name, bc in bcs:
        P = cfd.state2primitive(fluid, Q)
        impose_bc(lambda b, Pi: bc(Pi, b.normals), dom, name, P)
        g = np.unique(np.concatenate([b.ghost_indices for b in dom.boundaries[name].values()] or [np.zeros(0, np.int64)]))
        Q[g] = cfd.primitive2state(fluid, P[g])
        touched.append(g)
    return np.unique(np.concatenate(touched)) if touched else np.zeros(0, np.int64)
