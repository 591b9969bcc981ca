"""Host-side mirror of ``src/turbulence.jl``: pointwise closures on device vectors (``csrc/closures.cu``).

Velocity-gradient arguments are nested lists ``g[i][j]`` = d u_i / d x_j of device vectors, the reference's
matrix of vectors.  Named-tuple results of the reference come back as dicts with ASCII keys."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call
from .cfd import _grad_table, _like
from .domain import context

F32 = np.float32


def _wall_params(kappa=0.41, C_=4.9, A=19.0, beta=0.075, beta_star=0.09, D=4.2, A_plus=360.0, omega_fixed_point=0.5, n_iter=20):
    return _lib.WallParams(float(F32(kappa)), float(F32(C_)), float(F32(A)), float(F32(beta)), float(F32(beta_star)), float(F32(D)),
                           float(F32(A_plus)), float(F32(omega_fixed_point)), int(n_iter))


def wall_function(*args, kappa=0.41, C=4.9, A=19.0, beta=0.075, beta_star=0.09, D=4.2, A_plus=360.0,
                  omega_fixed_point=0.5, n_iter=20):
    """``wall_function(Rey)`` -> y_plus, u_plus, mu_plus, k_plus, dudy_plus (``src/turbulence.jl:27-72``) or
    ``wall_function(y, u, nu)`` -> u_tau, nu_t, k, omega, eps, dudn (``:74-98``)."""
    w = _wall_params(kappa, C, A, beta, beta_star, D, A_plus, omega_fixed_point, n_iter)
    if len(args) == 1:
        names = ("y_plus", "u_plus", "mu_plus", "k_plus", "dudy_plus")
        out = {k: _like(args[0]) for k in names}
        call("ibx_wall_function_rey", context(), w, args[0].h, *[out[k].h for k in names])
        return out
    y, u, nu = args
    names = ("u_tau", "nu_t", "k", "omega", "eps", "dudn")
    out = {k: _like(y) for k in names}
    call("ibx_wall_function", context(), w, y.h, u.h, nu.h, *[out[k].h for k in names])
    return out


def shear_rate(velocity_gradient):
    """``shear_rate`` = sqrt(2 S_ij S_ij) (``src/turbulence.jl:110-124``)."""
    nd, tab = _grad_table(velocity_gradient)
    out = _like(velocity_gradient[0][0])
    call("ibx_shear_rate", context(), nd, tab, out.h)
    return out


def Smagorinsky_nuSGS(Delta, S, Cs=0.17):
    """``Smagorinsky_νSGS`` (``src/turbulence.jl:134-137``)."""
    out = _like(S)
    call("ibx_smagorinsky", context(), Delta.h, S.h, C.c_float(float(F32(Cs))), out.h)
    return out


def standard_keps(k, eps, S, Cmu=0.09, sigma_k=1.0, sigma_eps=1.3, C1eps=1.44, C2eps=1.92):
    """``standard_kϵ`` (``src/turbulence.jl:175-194``) -> nu_k, nu_eps, Sk, Seps, nu_t."""
    names = ("nu_k", "nu_eps", "Sk", "Seps", "nu_t")
    out = {n: _like(k) for n in names}
    f = lambda x: C.c_float(float(F32(x)))
    call("ibx_standard_keps", context(), k.h, eps.h, S.h, f(Cmu), f(sigma_k), f(sigma_eps), f(C1eps), f(C2eps),
         *[out[n].h for n in names])
    return out


def Wray_Agarwal(R, S, gradR, gradS, sigma_R=0.72, C1=0.0829, kappa=0.41):
    """``Wray_Agarwal`` (``src/turbulence.jl:222-241``) -> nu_t (= R), nu_R, S; gradients are N x nd."""
    nuR, So = _like(R), _like(R)
    f = lambda x: C.c_float(float(F32(x)))
    call("ibx_wray_agarwal", context(), R.h, S.h, gradR.h, gradS.h, f(sigma_R), f(C1), f(kappa), nuR.h, So.h)
    return {"nu_t": R, "nu_R": nuR, "S": So}


def Ducros_sensor(velocity_gradient):
    """``Ducros_sensor`` (``src/turbulence.jl:253-283``)."""
    nd, tab = _grad_table(velocity_gradient)
    out = _like(velocity_gradient[0][0])
    call("ibx_ducros_sensor", context(), nd, tab, out.h)
    return out


def WALE_nuSGS(Delta, velocity_gradient, Cw=0.325):
    """``WALE_νSGS`` (``src/turbulence.jl:292-337``), 3-D only."""
    nd, tab = _grad_table(velocity_gradient)
    if nd != 3:
        raise ValueError("WALE model only implemented for 3D")
    out = _like(Delta)
    call("ibx_wale", context(), Delta.h, tab, C.c_float(float(F32(Cw))), out.h)
    return out
