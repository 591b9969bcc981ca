"""Oracle (test infrastructure) restatement of ``src/point_implicit.jl`` (module PointImplicit)."""
import numpy as np

from .nninterp import pinv


def hutchinson_trick(f, x, n_samples, h=1e-6, pre_evaluated_fx=None, rng=None, probes=None):
    """``hutchinson_trick``, ``src/point_implicit.jl:18-91``.

    Vector form returns the diagonal estimate; matrix form returns D[p, j, i] ~ d f_j / d x_i.
    ``probes`` (n_samples [, nv], npoints) of +-1 may be supplied to make the estimate
    reproducible (the reference draws from the global RNG).
    """
    x = np.asarray(x)
    if x.ndim == 2:
        fX = f(x) if pre_evaluated_fx is None else pre_evaluated_fx
        Xb = x.copy()
        cols = []
        for i in range(x.shape[1]):
            def fv(xi, i=i):
                Xb[:, i] = xi
                out = f(Xb)
                Xb[:, i] = x[:, i]
                return out
            pr = None if probes is None else probes[:, i]
            cols.append(hutchinson_trick(fv, x[:, i].copy(), n_samples, h, fX, rng, pr))
        return np.stack(cols, axis=-1)
    fx = f(x) if pre_evaluated_fx is None else pre_evaluated_fx
    s = np.zeros_like(fx)
    rng = rng or np.random.default_rng(0)
    for k in range(n_samples):
        z = probes[k] if probes is not None else rng.choice(np.array([-1, 1], dtype=np.int32), size=x.shape[0])
        Jz = (f(x + z * h) - fx) / h
        zz = z if s.ndim == 1 else z[:, None]
        s = s + zz * Jz
    return s / n_samples


class Linearization:
    """``Linearization``, ``src/point_implicit.jl:98-114``."""

    def __init__(self, f, x, fx, h):
        self.f, self.x, self.fx, self.h = f, x, fx, h

    def __call__(self, v):
        return (self.f(self.x + v * self.h) - self.fx) / self.h


def inverse_blocks(D):
    """``_inverse_blocks!``, ``src/point_implicit.jl:125-135``."""
    if D.ndim == 1:
        return D.dtype.type(1.0) / (np.finfo(D.dtype).eps + D)
    return pinv(D)


class PIPreconditioner:
    """``PIPreconditioner``, ``src/point_implicit.jl:121-161``: out[p, j] = sum_i Dinv[p, j, i] v[p, i]."""

    def __init__(self, inverse_diagonal):
        self.inverse_diagonal = inverse_diagonal

    def __call__(self, v):
        if self.inverse_diagonal.ndim == 1:
            return v * self.inverse_diagonal
        D = self.inverse_diagonal
        acc = v[:, None, 0] * D[:, :, 0]
        for i in range(1, D.shape[2]):
            acc = acc + v[:, None, i] * D[:, :, i]
        return acc


def linearize(f, x, n_hutchinson_samples=30, pre_evaluated_fx=None, h=1e-6, rng=None, probes=None):
    """``linearize``, ``src/point_implicit.jl:184-207`` -> (A, b, D)."""
    fx = f(x) if pre_evaluated_fx is None else pre_evaluated_fx.copy()
    x = x.copy()
    D = hutchinson_trick(f, x, n_hutchinson_samples, h, pre_evaluated_fx, rng, probes)
    return Linearization(f, x, fx, h), -fx, PIPreconditioner(inverse_blocks(D))


def proj_along(A, v, b):
    """``proj_along``, ``src/point_implicit.jl:220-233``."""
    eps = np.finfo(v.dtype).eps
    Av = A(v)
    return np.vdot(Av, b) / (np.vdot(Av, Av) + eps), Av


def solve(A, b, prec, n_iter=100, n_inner=1, rtol=1e-2, atol=1e-7, multigrid=None):
    """``solve``, ``src/point_implicit.jl:250-329`` -> (x, |r|/|r0|)."""
    eps = np.finfo(b.dtype).eps
    nr0 = np.linalg.norm(b)
    nr = nr0
    x = np.zeros_like(b)
    r = b.copy()
    n_levels = 0 if multigrid is None else len(multigrid.coarseners)
    n_mgrid = n_levels
    for _ in range(n_iter):
        for _ in range(n_inner):
            s = prec(r)
            if n_mgrid > 0:
                s = multigrid.prolongators[n_mgrid - 1](multigrid.coarseners[n_mgrid - 1](s))
            alpha, As = proj_along(A, s, r)
            x = x + s * alpha
            r = r - As * alpha
            s = r / (eps + np.max(np.abs(r)))
            alpha, As = proj_along(A, s, r)
            x = x + s * alpha
            r = r - As * alpha
            nr = np.linalg.norm(r)
            if nr < nr0 * rtol + atol:
                return x, nr / (nr0 + eps)
        n_mgrid = n_levels if n_mgrid == 0 else n_mgrid - 1
    return x, nr / (nr0 + eps)
