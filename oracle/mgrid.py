"""Oracle (test infrastructure) restatement of ``src/mgrid.jl`` (module GeometricMultigrid)."""
import numpy as np

from .accumulator import Accumulator
from .nninterp import KDTree


def coarsener_and_prolongator(X, n, volumes=None, perm=None):
    """``coarsener_and_prolongator``, ``src/mgrid.jl:24-97``; X (points, N), point index first.

    ``perm`` stands in for ``randperm`` when ``random_permutation = true`` (the
    reference's RNG stream is not reproducible here).
    """
    npts, N = X.shape
    if volumes is None:
        volumes = np.ones(npts, dtype=X.dtype)
    Xs = X if perm is None else X[perm]
    Xc = Xs[:: 2 ** (N * n)]
    idxs, _ = KDTree(Xc).nn(X)
    order = np.argsort(idxs, kind="stable")
    cnt = np.bincount(idxs, minlength=Xc.shape[0])
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    members = order  # fine points of every cluster, ascending
    v = volumes[members]
    tot = np.zeros(Xc.shape[0], dtype=volumes.dtype)
    # sum(v) over each cluster, sequential
    for c in range(Xc.shape[0]):
        s = volumes.dtype.type(0)
        for m in members[ptr[c]:ptr[c + 1]]:
            s = s + volumes[m]
        tot[c] = s
    w = v / np.repeat(tot, cnt)
    coarsener = Accumulator.from_csr(ptr, members, w, first_index=True)
    prolongator = Accumulator.from_csr(np.arange(npts + 1), idxs, None, first_index=True)
    return coarsener, prolongator


class Multigrid:
    """``Multigrid``, ``src/mgrid.jl:104-144``."""

    def __init__(self, X, n_levels, volumes=None):
        self.coarseners, self.prolongators = [], []
        for n in range(1, n_levels + 1):
            c, p = coarsener_and_prolongator(X, n, volumes)
            self.coarseners.append(c)
            self.prolongators.append(p)
