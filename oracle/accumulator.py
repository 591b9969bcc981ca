"""Oracle (test infrastructure) restatement of ``src/accumulator.jl``.

Indices are 0-based here; the reference is 1-based.  Buckets keep the reference's
layout: one ``(rows, stencil[l, n_l], weights[l, n_l] | None)`` tuple per stencil
length ``l`` (``src/accumulator.jl:44-61``).
"""
import numpy as np


class Accumulator:
    """``Accumulator`` struct and constructor, ``src/accumulator.jl:12-65``."""

    def __init__(self, inds, weights=None, first_index=False):
        n = len(inds)
        ls = np.fromiter((len(s) for s in inds), dtype=np.int64, count=n)
        self.n_output = n
        self.first_index = first_index
        self.stencils = {}
        # `unique(ls)` keeps first-appearance order (accumulator.jl:47); buckets write
        # disjoint rows so the order is immaterial.
        for l in dict.fromkeys(ls.tolist()):
            rows = np.flatnonzero(ls == l)
            if l == 0:
                st = np.zeros((0, rows.size), dtype=np.int64)
                ws = None if weights is None else np.zeros((0, rows.size), dtype=np.float32)
            else:
                st = np.stack([np.asarray(inds[r], dtype=np.int64) for r in rows], axis=1)
                ws = None
                if weights is not None:
                    ws = np.stack([np.asarray(weights[r]) for r in rows], axis=1)
            self.stencils[int(l)] = (rows, st, ws)

    @classmethod
    def from_csr(cls, ptr, idx, w=None, first_index=True):
        """Build from CSR arrays without Python lists (used by the big-table builders)."""
        self = cls.__new__(cls)
        ptr = np.asarray(ptr, dtype=np.int64)
        ls = np.diff(ptr)
        self.n_output = ls.size
        self.first_index = first_index
        self.stencils = {}
        for l in np.unique(ls):
            rows = np.flatnonzero(ls == l)
            gather = ptr[rows][None, :] + np.arange(l, dtype=np.int64)[:, None]
            st = np.asarray(idx, dtype=np.int64)[gather]
            ws = None if w is None else np.asarray(w)[gather]
            self.stencils[int(l)] = (rows, st, ws)
        return self

    def _call_vec(self, v, delta=False, f=None, op=None):
        """Vector call, ``src/accumulator.jl:78-111``.

        Products are formed first, then reduced top-to-bottom along the stencil
        (``reduce(op, ...; dims = 1)``), one rounding per operation, no FMA.
        """
        # the reference allocates `similar(v)`; we widen to the weight type so that the
        # docstring example (Int values, Float64 weights) prints [3.0, 38.0] as documented
        wdt = [ws.dtype for _, _, ws in self.stencils.values() if ws is not None]
        out = np.zeros(self.n_output, dtype=np.result_type(v.dtype, *wdt) if v.dtype.kind in "iu" else v.dtype)
        for rows, st, ws in self.stencils.values():
            l = st.shape[0]
            if l == 0 or rows.size == 0:
                continue  # reduce over a 0-row matrix leaves the zero initialisation
            vals = v[st]
            if ws is None:
                if f is not None:
                    vals = f(vals)
                terms = vals
            else:
                if delta:
                    vals = vals - v[rows][None, :]
                if f is not None:
                    vals = f(vals)
                terms = vals * ws
            acc = terms[0]
            for j in range(1, l):
                acc = (acc + terms[j]) if op is None else op(acc, terms[j])
            out[rows] = acc.astype(out.dtype, copy=False)
        return out

    def __call__(self, v, delta=False, f=None, op=None):
        """Array call, ``src/accumulator.jl:126-130`` (``mapslices`` over the summation axis)."""
        v = np.asarray(v)
        if v.ndim == 1:
            return self._call_vec(v, delta, f, op)
        axis = 0 if self.first_index else v.ndim - 1
        moved = np.moveaxis(v, axis, 0)
        flat = moved.reshape(moved.shape[0], -1)
        cols = [self._call_vec(np.ascontiguousarray(flat[:, c]), delta, f, op) for c in range(flat.shape[1])]
        res = np.stack(cols, axis=1).reshape((self.n_output,) + moved.shape[1:])
        return np.moveaxis(res, 0, axis)

    def decompose(self):
        """``decompose``, ``src/accumulator.jl:137-165``."""
        indices = [None] * self.n_output
        weights = [None] * self.n_output
        has_w = False
        for rows, st, ws in self.stencils.values():
            for k, i in enumerate(rows):
                indices[i] = st[:, k].copy()
                if ws is not None:
                    has_w = True
                    weights[i] = ws[:, k].copy()
        return (indices, weights) if has_w else indices

    def to_csr(self):
        """(ptr, idx, w|None) in row order; helper for comparing against device tables."""
        ls = np.zeros(self.n_output, dtype=np.int64)
        for l, (rows, _, _) in self.stencils.items():
            ls[rows] = l
        ptr = np.concatenate([[0], np.cumsum(ls)])
        idx = np.zeros(ptr[-1], dtype=np.int64)
        any_w = any(ws is not None for _, _, ws in self.stencils.values())
        w = np.zeros(ptr[-1], dtype=np.float32) if any_w else None
        for l, (rows, st, ws) in self.stencils.items():
            if l == 0:
                continue
            pos = ptr[rows][None, :] + np.arange(l)[:, None]
            idx[pos] = st
            if ws is not None:
                w[pos] = ws
        return ptr, idx, w

    def re_index(self, hmap):
        """``NNInterpolator.re_index!``, ``src/nninterp.jl:175-183`` (hmap: array old->new)."""
        for l, (rows, st, ws) in list(self.stencils.items()):
            self.stencils[l] = (rows, hmap[st], ws)

    def domain(self):
        """``NNInterpolator.domain``, ``src/nninterp.jl:147-168``: sorted unique stencil indices."""
        parts = [st.ravel() for _, st, _ in self.stencils.values() if st.size]
        if not parts:
            return np.zeros(0, dtype=np.int64)
        return np.unique(np.concatenate(parts))
