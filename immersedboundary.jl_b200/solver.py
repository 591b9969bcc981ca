"""Host-side mirrors of ``src/solver.jl`` (``FAS!``), ``src/mgrid.jl`` (``Multigrid``) and
``src/point_implicit.jl`` (``linearize``/``solve``) driving device arrays: host recursion and scalars,
all array work in libibx kernels."""
import ctypes as C

import numpy as np

from ._lib import call, ptr
from .domain import Accumulator, DeviceArray, context, dot

F32 = np.float32
EPS32 = np.finfo(np.float32).eps


def FAS(f, Q, coarseners=(), prolongators=(), prescribed_f=None, multigrid_level=0, n_iter=50, rtol=1e-1, atol=1e-7):
    """``FAS!`` (``src/solver.jl:39-91``) on ``DeviceArray``s; ``f(level, Q) -> (r, omega)``; mutates ``Q``.

    Keeps the reference's ``length(coarseners) > 1`` quirk (``:60``)."""
    l = multigrid_level
    fQ, omega = f(l, Q)
    source = None
    if prescribed_f is not None:
        source = prescribed_f - fQ
    r = fQ if source is None else fQ + source
    nr0 = float(r.norm())
    nr = nr0
    if len(coarseners) > 1:
        Qc = coarseners[0](Q)
        Qcold = Qc.copy()
        pfQc = coarseners[0](r)
        FAS(f, Qc, coarseners[1:], prolongators[1:], pfQc, l + 1, n_iter, rtol, atol)
        Q += prolongators[0](Qc - Qcold)
    for _ in range(n_iter):
        r, omega = f(l, Q)
        if source is not None:
            r += source
        om = omega if isinstance(omega, DeviceArray) else DeviceArray(Q.rows, 1, True).fill(omega)
        call("ibx_clamped_update", context(), Q.h, om.h, r.h)
        nr = float(r.norm())
        if nr < nr0 * rtol + atol:
            break
    return nr / (nr0 + EPS32)


RK_STAGES = {1: (1.0,), 2: (0.5, 1.0), 3: (0.1481, 0.4, 1.0), 4: (0.25, 1.0 / 3.0, 0.5, 1.0), 5: (0.25, 1.0 / 6.0, 0.375, 0.5, 1.0)}


def local_step_update(Q0, R, cfl, alpha, Q, mask=None):
    """``Q = Q0 + ((alpha / cfl) * R) * mask`` in one kernel (``ibx_local_step_update``); ``Q`` may be ``Q0``."""
    call("ibx_local_step_update", context(), Q0.h, R.h, cfl.h, 0 if mask is None else mask.h, float(F32(alpha)), Q.h)
    return Q


def march_euler(dom, fluid, bcs, Q, n_steps, CFL=0.8, stages=3, live=None, flux="hll", monitor=None, every=0, native=True, graph=True):
    """Pseudo-time march of the Euler residual with local time steps, the driver loop the reference leaves to its users
    (``test/advection.jl:28-46`` applies the BCs to the marched state and advances it by ``R * CFL / cfl``; ``FAS!``
    with ``f(l, Q) = (R .* CFL ./ cfl, 1)``, ``src/solver.jl:78-82``, is the single-stage case).

    Per step: ghost update of ``Q`` in place, ``Q0 = Q``, then for each multistage coefficient ``a``:
    ``step_euler`` (ghost update + residual) and ``Q = Q0 + a CFL R / cfl * live``.  ``live`` (N x 1 of 0/1) freezes
    cells between residual evaluations -- the ghost cells, which only ever take boundary values: a ghost cell advanced by
    its own residual drifts, and so does every image point that interpolates from it.  Everything runs on the device;
    ``monitor(step, Q, R, cfl)`` is called every ``every`` steps with the arrays of the last stage.

    Without a monitor the whole loop runs inside the library (``ibx_march_euler``, ``native=True``): one step is captured
    into a CUDA graph and replayed (``graph=True``) -- on a small mesh the loop is launch-bound -- with the same kernels
    in the same order, hence the same bits as the host loop below.  ``march_euler.last_graph_used`` tells whether the
    graph path ran."""
    from . import _lib, cfd

    alphas = RK_STAGES[stages] if isinstance(stages, int) else tuple(stages)
    if native and monitor is None and not getattr(dom, "shard_info", None):
        dom.upload()
        specs = (_lib.BCSpec * max(len(bcs), 1))(*[bc.spec(dom.boundary_index[name]) for name, bc in bcs])
        al = np.ascontiguousarray(alphas, dtype=F32)
        used = C.c_int(0)
        call("ibx_march_euler", context(), dom._h, fluid.c, 0 if flux == "hll" else 1, len(bcs), specs, Q.h, 0 if live is None else live.h,
             int(n_steps), float(F32(CFL)), len(al), ptr(al), int(bool(graph)), C.byref(used))
        march_euler.last_graph_used = bool(used.value)
        return Q
    R = DeviceArray(Q.rows, Q.cols, False)
    cf = DeviceArray(Q.rows, 1, True)
    Q0 = Q.like()
    for it in range(n_steps):
        cfd.ghost_update_euler(dom, fluid, Q, bcs)
        Q0.assign(Q)
        for a in alphas:
            cfd.step_euler(dom, fluid, bcs, Q, R, cf, flux)
            local_step_update(Q0, R, cf, float(F32(a) * F32(CFL)), Q, live)
        if monitor is not None and every and (it + 1) % every == 0:
            monitor(it + 1, Q, R, cf)
    return Q


class Multigrid:
    """``GeometricMultigrid.Multigrid(X, n_levels, volumes)`` (``src/mgrid.jl:104-144``)."""

    def __init__(self, X, n_levels, volumes=None):
        X = np.ascontiguousarray(X, dtype=F32)
        v = None if volumes is None else np.ascontiguousarray(volumes, dtype=F32)
        self.coarseners, self.prolongators = [], []
        for n in range(1, n_levels + 1):
            c, p = C.c_void_p(), C.c_void_p()
            call("ibx_mgrid_build", X.shape[1], X.shape[0], ptr(X), n, ptr(v), C.byref(c), C.byref(p))
            self.coarseners.append(Accumulator(_handle=c))
            self.prolongators.append(Accumulator(_handle=p))


class PIPreconditioner:
    """``PIPreconditioner`` (``src/point_implicit.jl:121-161``): per-cell block (or scalar) inverse."""

    def __init__(self, D, nv=None):
        self.nv = nv
        if nv is None:  # scalar diagonal: 1 / (eps + D)
            self.inv = (D + float(EPS32))._un(4)
        else:
            self.inv = DeviceArray(D.rows, nv * nv, False)
            call("ibx_block_pinv", context(), D.h, nv, self.inv.h)

    def __call__(self, v):
        if self.nv is None:
            return v * self.inv
        out = DeviceArray(v.rows, v.cols, v.vector)
        call("ibx_block_apply", context(), self.inv.h, self.nv, v.h, out.h)
        return out


def hutchinson_trick(f, X, n_samples, h=1e-6, fX=None, rng=None, probes=None):
    """``hutchinson_trick`` (``src/point_implicit.jl:18-91``) -> D (N x nv*nv, block (p, j, i) at column j + nv*i).

    Every probe is one full evaluation of ``f``.  ``probes`` (n_samples, nv, N) of +-1 makes the estimate reproducible
    and shard-independent (``synthetic.probe_signs``: counter-based, keyed by the global cell id); otherwise they come
    from a NumPy generator seeded by the caller (the reference uses the global RNG).  Device arrays are Float32, so
    unlike the reference -- whose Float64 default ``h = 1e-6`` promotes the whole evaluation to Float64
     -- ``h`` must be resolvable in Float32 next to ``X`` (scale the unknowns)."""
    rng = rng or np.random.default_rng(0)
    fX = f(X) if fX is None else fX
    N, nv = X.rows, X.cols
    D = DeviceArray(N, nv * nv, False).fill(0.0)
    for i in range(nv):
        acc = DeviceArray(N, nv, False).fill(0.0)
        for k in range(n_samples):
            z = np.ascontiguousarray(probes[k, i], dtype=F32) if probes is not None else rng.choice(np.array([-1.0, 1.0], dtype=F32), size=N)
            dz = DeviceArray.from_host(z)
            Xb = X.copy()
            col = X.col(i) + dz * float(h)
            call("ibx_array_set_column", context(), Xb.h, i, col.h)
            J = (f(Xb) - fX) / float(h)
            acc += J * dz
        acc = acc / float(n_samples)
        for j in range(nv):
            colj = acc.col(j)  # keep the temporary alive across the call
            call("ibx_array_set_column", context(), D.h, j + nv * i, colj.h)
    return D


class Linearization:
    """``Linearization`` (``src/point_implicit.jl:98-114``): finite-difference Jacobian-vector product."""

    def __init__(self, f, x, fx, h):
        self.f, self.x, self.fx, self.h = f, x, fx, float(h)

    def __call__(self, v):
        return (self.f(self.x + v * self.h) - self.fx) / self.h


def linearize(f, x, n_hutchinson_samples=30, pre_evaluated_fx=None, h=1e-6, rng=None, probes=None):
    """``linearize`` (``src/point_implicit.jl:184-207``) -> (A, b, D)."""
    fx = f(x) if pre_evaluated_fx is None else pre_evaluated_fx.copy()
    x = x.copy()
    D = hutchinson_trick(f, x, n_hutchinson_samples, h, fx, rng, probes)
    return Linearization(f, x, fx, h), -fx, PIPreconditioner(D, x.cols)


def proj_along(A, v, b):
    """``proj_along`` (``src/point_implicit.jl:220-233``)."""
    Av = A(v)
    return dot(Av, b) / (dot(Av, Av) + float(EPS32)), Av


def solve(A, b, prec, n_iter=100, n_inner=1, rtol=1e-2, atol=1e-7, multigrid=None):
    """``solve`` (``src/point_implicit.jl:250-329``) -> (x, |r|/|r0|)."""
    eps = float(EPS32)
    nr0 = float(b.norm())
    nr = nr0
    x = b.like().fill(0.0)
    r = b.copy()
    n_levels = 0 if multigrid is None else len(multigrid.coarseners)
    n_mgrid = n_levels
    for _ in range(n_iter):
        for _ in range(n_inner):
            s = prec(r)
            if n_mgrid > 0:
                s = multigrid.prolongators[n_mgrid - 1](multigrid.coarseners[n_mgrid - 1](s))
            alpha, As = proj_along(A, s, r)
            x += s * alpha
            r -= As * alpha
            s = r / (eps + float(r.maxabs()))
            alpha, As = proj_along(A, s, r)
            x += s * alpha
            r -= As * alpha
            nr = float(r.norm())
            if nr < nr0 * rtol + atol:
                return x, nr / (nr0 + eps)
        n_mgrid = n_levels if n_mgrid == 0 else n_mgrid - 1
    return x, nr / (nr0 + eps)
