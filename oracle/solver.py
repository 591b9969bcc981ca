"""Oracle (test infrastructure) restatement of ``src/solver.jl`` (``FAS!``)."""
import numpy as np

F32 = np.float32


def _norm(r):
    return F32(np.sqrt(np.sum(np.asarray(r, dtype=np.float64) ** 2)))


def FAS(f, Q, coarseners=(), prolongators=(), prescribed_f=None, multigrid_level=0, n_iter=50, rtol=F32(1e-1),
        atol=F32(1e-7)):
    """``FAS!``, ``src/solver.jl:39-91``; mutates ``Q`` in place, returns the norm-reduction ratio.

    Quirk kept from the reference: the coarse recursion only happens when
    ``length(coarseners) > 1`` (``:60``), so a single-level hierarchy is never used.
    """
    l = multigrid_level
    fQ, omega = f(l, Q)
    source = F32(0.0)
    if prescribed_f is not None:
        source = prescribed_f - fQ
    r = fQ + source
    nr0 = _norm(r)
    nr = nr0
    if len(coarseners) > 1:
        coars, prolong = coarseners[0], prolongators[0]
        Qc = coars(Q)
        Qcold = Qc.copy()
        pfQc = coars(r)
        FAS(f, Qc, coarseners=coarseners[1:], prolongators=prolongators[1:], prescribed_f=pfQc,
            multigrid_level=l + 1, n_iter=n_iter, atol=atol, rtol=rtol)
        Q += prolong(Qc - Qcold)
    for _ in range(n_iter):
        r, omega = f(l, Q)
        r = r + source
        Q += np.clip(omega, F32(0.0), F32(1.0)) * r
        nr = _norm(r)
        if nr < nr0 * rtol + atol:
            break
    return nr / (nr0 + np.finfo(np.float32).eps)
