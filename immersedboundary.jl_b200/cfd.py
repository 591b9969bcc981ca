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

    @property
    def transport(self):
        """Sutherland / conductivity constants as the ABI's ``ibx_transport``."""
        k = (C.c_float * 4)(*([float(x) for x in self.k] + [0.0] * (4 - len(self.k))))
        return _lib.Transport(float(self.mu_ref), float(self.T_ref), float(self.S), len(self.k), k)


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


def dynamic_viscosity(fluid, T):
    """``dynamic_viscosity`` (Sutherland, ``src/cfd.jl:71-77``)."""
    mu = _like(T)
    call("ibx_dynamic_viscosity", context(), fluid.transport, T.h, mu.h)
    return mu


def heat_conductivity(fluid, T):
    """``heat_conductivity`` (``src/cfd.jl:84-90``)."""
    k = _like(T)
    call("ibx_heat_conductivity", context(), fluid.transport, T.h, k.h)
    return k


def _handles(arrays):
    return (C.c_int64 * len(arrays))(*[a.h for a in arrays])


def _grad_table(g):
    """nested ``g[i][j]`` = d u_i / d x_j (vectors) -> row-major handle table."""
    nd = len(g)
    assert all(len(r) == nd for r in g), "velocity-gradient matrix must be nd x nd"
    return nd, _handles([g[i][j] for i in range(nd) for j in range(nd)])


def viscous_fluxes(fluid, P, Pgrad, dim, mu_t=0.0):
    """``viscous_fluxes(fluid, P, Pgrad, dim; μₜ)`` (``src/cfd.jl:664-736``): ``dim`` a 0-based axis or an N x nd
    direction matrix; ``mu_t`` a scalar or a vector."""
    F = _like(P)
    normals = dim if isinstance(dim, DeviceArray) else None
    mt = mu_t if isinstance(mu_t, DeviceArray) else None
    call("ibx_viscous_fluxes", context(), fluid.transport, P.h, _handles(list(Pgrad)), -1 if normals is not None else int(dim),
         normals.h if normals is not None else 0, mt.h if mt is not None else 0, C.c_float(0.0 if mt is not None else float(mu_t)), F.h)
    return F


def JST_sensor_3pt(Pim1, Pi, Pip1):
    """Pointwise ``JST_sensor(Pim1, Pi, Pip1)`` (``src/cfd.jl:563-573``)."""
    out = _like(Pi)
    call("ibx_jst_sensor3", context(), Pim1.h, Pi.h, Pip1.h, out.h)
    return out


def shock_sensor(velocity_gradients):
    """``shock_sensor`` (``src/cfd.jl:589-617``); ``velocity_gradients[i][j]`` = d u_i / d x_j."""
    nd, tab = _grad_table(velocity_gradients)
    out = _like(velocity_gradients[0][0])
    call("ibx_shock_sensor", context(), nd, tab, out.h)
    return out


def pressure_coefficient(fluid, p, p_inf, M_inf):
    """``pressure_coefficient`` (``src/cfd.jl:420-426``)."""
    Cp = _like(p)
    call("ibx_pressure_coefficient", context(), C.c_float(float(fluid.gamma)), p.h, C.c_float(float(p_inf)), C.c_float(float(M_inf)), Cp.h)
    return Cp


def streamwise_direction(alpha, beta=None):
    """``streamwise_direction`` (``src/cfd.jl:399-409, 434-436``), angles in degrees."""
    ca, sa = np.cos(np.radians(alpha)), np.sin(np.radians(alpha))
    if beta is None:
        return np.array([ca, sa])
    cb, sb = np.cos(np.radians(beta)), np.sin(np.radians(beta))
    return np.array([ca * cb, -ca * sb, sa])


def Reynolds_number(fluid, P_inf, L_ref):
    """``Reynolds_number`` (``src/cfd.jl:626-637``); host scalars, the viscosity through the device kernel."""
    P_inf = np.asarray(P_inf, dtype=F32)
    V = F32(np.sqrt(np.sum(P_inf[2:].astype(np.float64) ** 2)))
    rho = P_inf[0] / (fluid.R * P_inf[1])
    mu = dynamic_viscosity(fluid, DeviceArray.from_host(P_inf[1:2].copy())).to_host()[0]
    return V * F32(L_ref) * rho / mu


def adjust_Reynolds(fluid, P_inf, L_ref, Re):
    """``adjust_Reynolds`` (``src/cfd.jl:645-654``): a new ``Fluid`` whose reference viscosity gives ``Re``."""
    mu_ref = fluid.mu_ref * Reynolds_number(fluid, P_inf, L_ref) / F32(Re)
    return Fluid(fluid.R, fluid.gamma, fluid.k, mu_ref, fluid.T_ref, fluid.S)


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

    def __call__(self, P, normals, image_distances=None, du_dn=None, transpiration=0.0):
        """``bc(P, normals; image_distances, du!dn, transpiration)`` (``src/cfd.jl:243-300``)."""
        out = _like(P)
        if image_distances is None and du_dn is None and not isinstance(transpiration, DeviceArray) and transpiration == 0.0:
            call("ibx_flowbc", context(), self.fluid.c, ptr(self.P), len(self.P), int(self.normal_flow), P.h, normals.h, out.h)
            return out
        tr = transpiration if isinstance(transpiration, DeviceArray) else None
        call("ibx_flowbc_ex", context(), self.fluid.c, ptr(self.P), len(self.P), int(self.normal_flow), P.h, normals.h,
             image_distances.h if image_distances is not None else 0, du_dn.h if du_dn is not None else 0,
             tr.h if tr is not None else 0, C.c_float(0.0 if tr is not None else float(transpiration)), out.h)
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


def residual_rans(dom, fluid, Q, qR, R, RR, cfl, sigma_R=0.72, C1=0.0829, kappa=0.41):
    """Canonical RANS residual of configuration C5 (``ibx_residual_rans``): ``R, cfl`` = Euler residual + viscous fluxes
    with the Wray-Agarwal eddy viscosity, ``RR`` = residual of the transported ``qR = rho R`` (``src/cfd.jl:664-736``,
    ``src/turbulence.jl:197-241``; composition: ``oracle/euler.py: rans_residual``)."""
    dom.upload()
    call("ibx_residual_rans", context(), dom._h, fluid.c, fluid.transport, float(sigma_R), float(C1), float(kappa),
         Q.h, qR.h, R.h, RR.h, cfl.h)


def ghost_update_rans(dom, Q, qR, R_bcs):
    """IB ghost update of the transported Wray-Agarwal variable for ``R_bcs = [(boundary name, R value), ...]`` in order
    (``ibx_ghost_update_rans``); call after ``ghost_update_euler`` (it uses the ghost densities).  On a rank-local shard
    exchange ``Q`` in between: a donor that is a ghost cell of another rank must carry its new density too
    (``tools/mgpu_check.py`` checks the sequence against the single-domain result)."""
    dom.upload()
    coupled = (getattr(dom, "shard_info", None) or {}).get("coupled_families")
    done = []
    for name, val in R_bcs:
        if coupled and any((e, name) in coupled for e in done):
            dom.halo_exchange(qR)      # like ghost_update_euler: a ghost of this family reads a ghost of an earlier one on another rank
            done = []
        call("ibx_ghost_update_rans", context(), dom._h, dom.boundary_index[name], Q.h, qR.h, float(val))
        done.append(name)


def ghost_update_euler(dom, fluid, Q, bcs):
    """IB ghost update of the conservative state for ``bcs = [(boundary name, FlowBC), ...]`` in order."""
    dom.upload()
    coupled = (getattr(dom, "shard_info", None) or {}).get("coupled_families")
    done = []
    for name, bc in bcs:
        if coupled and any((e, name) in coupled for e in done):
            # a ghost of this family reads, on another rank, a ghost of a family applied above: refresh the halo rows
            # so that the sequence equals the reference's successive impose_bc! calls on one array
            dom.halo_exchange(Q)
            done = []
        call("ibx_ghost_update_euler", context(), dom._h, dom.boundary_index[name], fluid.c, ptr(bc.P), len(bc.P),
             int(bc.normal_flow), Q.h)
        done.append(name)


def step_euler(dom, fluid, bcs, Q, R, cfl, flux="hll"):
    """One step of a solver loop on a whole domain (``ibx_step_euler``): ghost updates of ``bcs`` then the residual, the ghost
    update hidden behind the residual of the blocks that do not read a ghost cell.  Same bits as ``ghost_update_euler`` +
    ``residual_euler``."""
    dom.upload()
    specs = (_lib.BCSpec * max(len(bcs), 1))(*[bc.spec(dom.boundary_index[name]) for name, bc in bcs])
    call("ibx_step_euler", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, len(bcs), specs, Q.h, R.h, cfl.h)


def step_euler_sharded(dom, fluid, bcs, Q, R, cfl, flux="hll"):
    """One step of a sharded solver loop (``ibx_step_euler_sharded``): halo exchange -> ghost updates of ``bcs`` -> halo
    exchange on the halo stream, hidden behind the residual of the blocks that read neither a ghost nor a halo cell; then
    the rest.  Same bits as ``halo_exchange; ghost_update_euler; halo_begin; residual_euler``."""
    dom.upload()
    specs = (_lib.BCSpec * max(len(bcs), 1))(*[bc.spec(dom.boundary_index[name]) for name, bc in bcs])
    coupled = (getattr(dom, "shard_info", None) or {}).get("coupled_families") or ()
    order = [name for name, _ in bcs]
    between = any((order[i], order[j]) in coupled for j in range(len(order)) for i in range(j))
    call("ibx_step_euler_sharded", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, len(bcs), specs, int(between), Q.h, R.h, cfl.h)


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
