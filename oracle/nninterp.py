"""Oracle (test infrastructure) restatement of ``src/nninterp.jl`` and of the
NearestNeighbors.jl (^0.4.21, not vendored) queries the reference relies on.

Tie rule (ours, documented in SURVEY.md 8c): candidates are ranked by
(squared Euclidean distance accumulated dimension by dimension in the promoted
float type, then lower index).  NearestNeighbors' own tie order depends on tree
traversal and cannot be reproduced without its source.
"""
import numpy as np
from scipy.spatial import cKDTree

from .accumulator import Accumulator


def _d2(points, x):
    """Sequential squared distance in the promoted dtype; points (m, nd) or (q, m, nd)."""
    dt = np.result_type(points.dtype, x.dtype)
    diff = points.astype(dt, copy=False) - x.astype(dt, copy=False)
    acc = diff[..., 0] * diff[..., 0]
    for d in range(1, diff.shape[-1]):
        acc = acc + diff[..., d] * diff[..., d]
    return acc


class KDTree:
    """Stand-in for ``NearestNeighbors.KDTree`` over the rows of ``data`` (n, nd)."""

    def __init__(self, data, leafsize=10):
        self.data = np.ascontiguousarray(data)
        self.n, self.nd = self.data.shape
        self._tree = cKDTree(self.data.astype(np.float64), leafsize=max(int(leafsize), 1))

    def knn(self, xq, k):
        """k nearest rows for each query row -> (idx (q,k), dist (q,k)), sorted by (d2, idx)."""
        xq = np.atleast_2d(np.asarray(xq))
        q = xq.shape[0]
        k = int(k)
        assert k <= self.n, "knn: k larger than the number of points"
        kk = min(self.n, k + 8)
        dd, ii = self._tree.query(xq.astype(np.float64), k=kk)
        if kk == 1:
            dd = dd[:, None]
            ii = ii[:, None]
        d2 = _d2(self.data[ii], xq[:, None, :])
        order = np.lexsort((ii, d2), axis=1)
        ii_s = np.take_along_axis(ii, order, axis=1)[:, :k]
        d2_s = np.take_along_axis(d2, order, axis=1)[:, :k]
        if kk < self.n:
            # tie sets that may have been truncated by the candidate window: redo exactly
            risky = np.flatnonzero(dd[:, kk - 1] <= dd[:, k - 1] * (1 + 1e-6) + 1e-300)
            for r in risky:
                cand = np.asarray(self._tree.query_ball_point(xq[r].astype(np.float64), dd[r, k - 1] * (1 + 1e-6) + 1e-300), dtype=np.int64)
                cd2 = _d2(self.data[cand], xq[r][None, :])
                o = np.lexsort((cand, cd2))[:k]
                ii_s[r] = cand[o]
                d2_s[r] = cd2[o]
        return ii_s.astype(np.int64), np.sqrt(d2_s)

    def nn(self, xq):
        i, d = self.knn(xq, 1)
        return i[:, 0], d[:, 0]

    def inrange(self, x, r):
        """Indices with distance <= r (boundary inclusive), ascending."""
        x = np.asarray(x)
        cand = np.asarray(self._tree.query_ball_point(x.astype(np.float64), float(r) * (1 + 1e-6) + 1e-300), dtype=np.int64)
        if cand.size == 0:
            return cand
        cd2 = _d2(self.data[cand], x[None, :])
        dt = cd2.dtype.type
        keep = cd2 <= dt(r) * dt(r)
        return np.sort(cand[keep])

    def inrange_many(self, xq, r):
        xq = np.asarray(xq)
        r = np.broadcast_to(np.asarray(r, dtype=np.float64), (xq.shape[0],))
        return self._tree.query_ball_point(xq.astype(np.float64), r * (1 + 1e-6) + 1e-300)


def pinv(A):
    """``LinearAlgebra.pinv`` with Julia's default ``rtol = eps(T) * min(m, n)``."""
    A = np.asarray(A)
    rtol = np.finfo(A.dtype).eps * min(A.shape[-2:])
    return np.linalg.pinv(A, rcond=rtol)


def _distances(dX):
    """``sum(dX .^ 2; dims = 1) |> sqrt .+ eps``, ``src/nninterp.jl:26-28`` (dX: (q, k, nd))."""
    acc = dX[..., 0] ** 2
    for d in range(1, dX.shape[-1]):
        acc = acc + dX[..., d] ** 2
    return np.sqrt(acc) + np.finfo(dX.dtype).eps


def linear_weights(X, idx, x):
    """``linear_weights``, ``src/nninterp.jl:16-42``, batched.

    X (n, nd); idx (q, k); x (q, nd) -> (w (q, k), mask (q, k)).
    """
    Tf = X.dtype
    eps = np.finfo(Tf).eps
    dX = X[idx] - x[:, None, :]
    dist = _distances(dX)
    w = (Tf.type(1.0) / dist).astype(Tf)
    A = np.concatenate([dX, np.ones(dX.shape[:2] + (1,), dtype=Tf)], axis=2)  # (q, k, nd+1)
    Aw = A * w[:, :, None]
    P = pinv(Aw)  # (q, nd+1, k)
    wts = P[:, -1, :] * w
    return wts.astype(Tf), np.abs(wts) > eps


def idw_weights(X, idx, x):
    """``IDW_weights``, ``src/nninterp.jl:47-69``, batched."""
    Tf = X.dtype
    eps = np.finfo(Tf).eps
    dX = X[idx] - x[:, None, :]
    dist = _distances(dX)
    w = (Tf.type(1.0) / dist).astype(Tf)
    s = w[:, 0].copy()
    for j in range(1, w.shape[1]):
        s = s + w[:, j]
    w = w / s[:, None]
    return w, np.abs(w) > np.sqrt(eps)


def interpolator_tables(X, Xc, tree=None, bias=None, linear=True, k=0):
    """Stencil/weight tables of ``Interpolator``, ``src/nninterp.jl:85-138``.

    X (n, nd) sources, Xc (q, nd) targets (point index first).  Returns
    (idx (q, k), w (q, k), mask (q, k)); entries with mask False are dropped by the
    reference (variable stencil length).
    """
    X = np.ascontiguousarray(X)
    Xc = np.ascontiguousarray(Xc)
    nd = X.shape[1]
    if k == 0:
        k = 2 ** nd
    if tree is None:
        tree = KDTree(X)
    Xq = Xc if bias is None else Xc + bias
    idx, _ = tree.knn(Xq, k)
    if linear:
        w, mask = linear_weights(X, idx, Xc)
    else:
        w, mask = idw_weights(X, idx, Xc)
    return idx, w, mask


def Interpolator(X, Xc, tree=None, bias=None, first_index=True, linear=True, k=0):
    """``Interpolator`` -> ``Accumulator``.  Arrays are always (points, nd) here, so
    ``first_index`` only selects the summation axis of the returned accumulator."""
    idx, w, mask = interpolator_tables(X, Xc, tree, bias, linear, k)
    lens = mask.sum(axis=1)
    ptr = np.concatenate([[0], np.cumsum(lens)])
    return Accumulator.from_csr(ptr, idx[mask], w[mask], first_index=first_index)
