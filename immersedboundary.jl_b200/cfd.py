"""Host-side mirror of the hot-path subset of ``src/cfd.jl`` on device arrays, plus the fused entry points."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, ptr
from .domain import DeviceArray, context

F32 = np.float32


class Fluid:
    """``Fluid`` (``src/cfd.jl:14-53``); only R and gamma enter the inviscid path."""

    def __init__(self, R=283.0, gamma=1.4, k=(0.00646, 6.468e-5), mu_ref=1.716e-5, T_ref=273.15, S=110.4):
        self.R, self.gamma = F32(R), F32(gamma)
        self.k, self.mu_ref, self.T_ref, self.S = [F32(x) for x in np.atleast_1d(k)], F32(mu_ref), F32(T_ref), F32(S)

    @property
    def c(self):
        return _lib.Fluid(float(self.R), float(self.gamma))


def _like(a):
    return DeviceArray(a.rows, a.cols, a.vector)


def state2primitive(fluid, Q):
    """``state2primitive`` (``src/cfd.jl:137-151``)."""
    P = _like(Q)
    call("ibx_state2primitive", context(), fluid.c, Q.h, P.h)
    return P


def primitive2state(fluid, P):
    """``primitive2state`` (``src/cfd.jl:106-123``)."""
    Q = _like(P)
    call("ibx_primitive2state", context(), fluid.c, P.h, Q.h)
    return Q


def speed_of_sound(fluid, T):
    """``speed_of_sound`` (``src/cfd.jl:62-64``)."""
    a = _like(T)
    call("ibx_speed_of_sound", context(), fluid.c, T.h, a.h)
    return a


def inviscid_fluxes(fluid, PL, PR, *args):
    """``inviscid_fluxes(fluid, PL, PR, dim)`` (HLL, ``src/cfd.jl:459-508``) or
    ``inviscid_fluxes(fluid, PL, PR, nuL, nuR, dim)`` (sensor-Rusanov, ``:516-554``); ``dim`` 0-based."""
    if len(args) == 1:
        F = DeviceArray(PL.rows, PL.cols, PL.vector, f64=True)  # the reference returns Float64 here (src/cfd.jl:504-507)
        call("ibx_inviscid_fluxes_hll", context(), fluid.c, PL.h, PR.h, int(args[0]), F.h)
    else:
        F = _like(PL)
        nuL, nuR, dim = args
        call("ibx_inviscid_fluxes_sensor", context(), fluid.c, PL.h, PR.h, nuL.h, nuR.h, int(dim), F.h)
    return F


class FlowBC:
    """``FlowBC(fluid, P; normal_flow)`` (``src/cfd.jl:160-300``); call with ``(P, normals)``."""

    def __init__(self, fluid, P, normal_flow=False):
        self.fluid = fluid
        self.P = np.ascontiguousarray(P, dtype=F32)
        self.normal_flow = bool(normal_flow)

    def __call__(self, P, normals):
        out = _like(P)
        call("ibx_flowbc", context(), self.fluid.c, ptr(self.P), len(self.P), int(self.normal_flow), P.h, normals.h, out.h)
        return out

    def spec(self, boundary_index):
        s = _lib.BCSpec(boundary=boundary_index, normal_flow=int(self.normal_flow), n_pinf=len(self.P))
        for i, v in enumerate(self.P):
            s.Pinf[i] = float(v)
        return s


# --------------------------------------------------------------------------------- fused entry points
def residual_euler(dom, fluid, Q, R, cfl, flux="hll"):
    """Canonical Euler residual on the whole (rank-local) domain, block-structured kernels."""
    dom.upload()
    call("ibx_residual_euler", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, Q.h, R.h, cfl.h)


def ghost_update_euler(dom, fluid, Q, bcs):
    """IB ghost update of the conservative state for ``bcs = [(boundary name, FlowBC), ...]`` in order."""
    dom.upload()
    for name, bc in bcs:
        call("ibx_ghost_update_euler", context(), dom._h, dom.boundary_index[name], fluid.c, ptr(bc.P), len(bc.P),
             int(bc.normal_flow), Q.h)


def residual_advection(dom, u, Cvel, ud, spec):
    """Linear-advection residual of ``test/advection.jl:67-83`` + CFL denominator, whole domain."""
    dom.upload()
    call("ibx_residual_advection", context(), dom._h, u.h, Cvel.h, ud.h, spec.h)


def euler_step_host(dom, fluid, bcs, Q_host, R_host, cfl_host, flux="hll"):
    """End-to-end call with HOST buffers (column-major float32): H2D(Q), ghost updates, residual, D2H(R, cfl)."""
    dom.upload()
    specs = (_lib.BCSpec * max(len(bcs), 1))(*[bc.spec(dom.boundary_index[name]) for name, bc in bcs])
    assert Q_host.flags.f_contiguous and R_host.flags.f_contiguous and Q_host.dtype == np.float32
    call("ibx_euler_step_host", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, len(bcs), specs,
         ptr(Q_host), ptr(R_host), ptr(cfl_host))


def euler_step_host_begin(dom, fluid, bcs, Q_host, R_host, cfl_host, slot, flux="hll"):
    """Asynchronous half of ``euler_step_host``: enqueue H2D(Q) -> ghost updates -> residual -> D2H(R, cfl) on
    ``slot`` (0 or 1) and return.  Independent evaluations on alternating slots overlap their copies (full-duplex
    PCIe) and compute; the host buffers must stay alive until ``euler_step_host_end(slot)``."""
    dom.upload()
    specs = (_lib.BCSpec * max(len(bcs), 1))(*[bc.spec(dom.boundary_index[name]) for name, bc in bcs])
    assert Q_host.flags.f_contiguous and R_host.flags.f_contiguous and Q_host.dtype == np.float32
    call("ibx_euler_step_host_begin", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, len(bcs), specs,
         ptr(Q_host), ptr(R_host), ptr(cfl_host), int(slot))


def euler_step_host_end(slot):
    """Block until the evaluation enqueued on ``slot`` has delivered R and cfl to its host buffers."""
    call("ibx_euler_step_host_end", context(), int(slot))


def pinned_empty(shape, order="F"):
    """float32 array backed by pinned host memory (``ibx_host_alloc``) for the end-to-end path."""
    n = int(np.prod(shape))
    p = C.c_void_p()
    call("ibx_host_alloc", n * 4, C.byref(p))
    buf = (C.c_float * n).from_address(p.value)
    a = np.frombuffer(buf, dtype=F32).reshape(shape, order=order)
    return a
