"""Oracle (test infrastructure): the named residual compositions the fused kernels are
checked against, written with the restated reference operators only.

* ``advection_residual`` / ``advection_spectral`` follow ``test/advection.jl:52-83``.
* ``dissipation_residual`` / ``dissipation_spectral`` follow ``test/dissipation.jl:54-77``.
* ``euler_residual`` / ``euler_ghost_update`` are OUR canonical Euler composition
  (SURVEY.md Appendix A.10 -- the reference ships no Euler residual); they mirror the
  structure of ``test/advection.jl:67-83`` with the ``src/cfd.jl`` fluxes and ``FlowBC``.
"""
import numpy as np

from . import cfd
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
