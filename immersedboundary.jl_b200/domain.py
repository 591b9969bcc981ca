"""Host-side mirror of the reference's Domain runtime (``src/ImmersedBoundary.jl``) on top of the C ABI.

``Domain(msh; ...)``, the partition callable ``dom(f, *args)``, ``impose_bc``, ``multigrid``,
``volume_integral`` and the grid operators keep the reference's names and argument meaning; the
arrays the user closure receives are ``DeviceArray`` handles (device-resident float32,
column-major), and all tables live on the GPU from ``Domain`` creation on.  Indices and ``dim``
arguments are 0-based here (the Julia wrapper converts at the ``ccall``).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, ptr, IbxError
from .mesher import Mesh

F32 = np.float32
I32 = np.int32

_ctx = None


def context(device=None):
    """The process-wide device context (``ibx_init``); created on first use."""
    global _ctx
    if _ctx is None:
        import os
        dev = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        h = C.c_void_p()
        call("ibx_init", dev, C.byref(h))
        _ctx = h
    return _ctx


def synchronize():
    call("ibx_sync", context())


def launch_count():
    n = C.c_int64()
    call("ibx_launch_count", context(), C.byref(n))
    return n.value


def set_option(name, value):
    """``ibx_set_option``: run-time options of the fused Euler residual -- ``"arithmetic"`` (0 reference-exact, 1 fast),
    ``"path"`` (0 marching kernels, 1 tile kernels, 2 gather kernels), ``"sensor"`` (1 JST blend, 0 ``D = nothing``)."""
    call("ibx_set_option", context(), name.encode(), int(value))


def get_option(name):
    v = C.c_int()
    call("ibx_get_option", context(), name.encode(), C.byref(v))
    return v.value


class options:
    """``with ib.options(path=1): ...`` -- set options for a block and restore them afterwards."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


# --------------------------------------------------------------------------------- device arrays
class DeviceArray:
    """Opaque float32 device array (rows x cols, column-major): what ``conv_to_backend`` would produce.

    Arithmetic operators dispatch to the library's elementwise kernels so that user closures written
    like the reference's (``ud -= green_gauss(part, (uL + uR) * Cf / 2 + abs(Cf) * (uL - uR) / 2, dim)``)
    run on the device unchanged.
    """
    __array_priority__ = 1000

    def __init__(self, rows, cols=1, vector=None, _handle=None, f64=False):
        self.rows, self.cols = int(rows), int(cols)
        self.vector = (self.cols == 1) if vector is None else vector  # 1-D (N,) vs (N, 1)
        self.f64 = bool(f64)  # Float64 payload: the HLL flux and what is derived from it (src/cfd.jl:504-507)
        if _handle is None:
            h = C.c_int64()
            call("ibx_array_alloc_f64" if f64 else "ibx_array_alloc", context(), self.rows, self.cols, C.byref(h))
            _handle = h.value
        self.h = _handle

    def __del__(self):
        try:
            if _ctx is not None and self.h:
                _lib.lib.ibx_array_free(_ctx, self.h)
        except Exception:
            pass

    @classmethod
    def from_host(cls, a):
        a = np.asarray(a, dtype=F32)
        out = cls(a.shape[0], 1 if a.ndim == 1 else a.shape[1], vector=a.ndim == 1)
        out.upload(a)
        return out

    def upload(self, a):
        a = np.asfortranarray(np.asarray(a, dtype=F32).reshape(self.rows, self.cols))
        call("ibx_array_upload", context(), self.h, ptr(a))
        return self

    def to_host(self):
        out = np.zeros((self.rows, self.cols), dtype=np.float64 if self.f64 else F32, order="F")
        call("ibx_array_download_f64" if self.f64 else "ibx_array_download", context(), self.h, ptr(out))
        return out[:, 0].copy() if self.vector else out

    @property
    def shape(self):
        return (self.rows,) if self.vector else (self.rows, self.cols)

    def like(self, cols=None, vector=None, f64=None):
        return DeviceArray(self.rows, self.cols if cols is None else cols, self.vector if vector is None else vector,
                           f64=self.f64 if f64 is None else f64)

    def copy(self):
        out = self.like()
        call("ibx_array_copy", context(), out.h, self.h)
        return out

    def fill(self, v):
        call("ibx_array_fill", context(), self.h, float(v))
        return self

    def assign(self, other):
        """``a .= other`` (array or scalar)."""
        if isinstance(other, DeviceArray):
            call("ibx_array_copy", context(), self.h, other.h)
        else:
            self.fill(other)
        return self

    # -- elementwise glue (ibx_ew_*)
    def _bin(self, op, other, reverse=False, out=None):
        if isinstance(other, DeviceArray):
            a, b = (other, self) if reverse else (self, other)
            big = a if a.cols >= b.cols else b
            out = out or big.like(f64=a.f64 or b.f64)  # Julia promotion; in-place forms keep the target type
            call("ibx_ew_binary", context(), op, a.h, b.h, out.h)
            return out
        out = out or self.like()
        call("ibx_ew_scalar", context(), op, self.h, float(other), int(reverse), out.h)
        return out

    def __add__(self, o): return self._bin(0, o)
    def __radd__(self, o): return self._bin(0, o, True)
    def __sub__(self, o): return self._bin(1, o)
    def __rsub__(self, o): return self._bin(1, o, True)
    def __mul__(self, o): return self._bin(2, o)
    def __rmul__(self, o): return self._bin(2, o, True)
    def __truediv__(self, o): return self._bin(3, o)
    def __rtruediv__(self, o): return self._bin(3, o, True)
    def __iadd__(self, o): return self._bin(0, o, out=self)
    def __isub__(self, o): return self._bin(1, o, out=self)
    def __imul__(self, o): return self._bin(2, o, out=self)
    def __itruediv__(self, o): return self._bin(3, o, out=self)

    def _un(self, op):
        out = self.like()
        call("ibx_ew_unary", context(), op, self.h, out.h)
        return out

    def __abs__(self): return self._un(0)
    def __neg__(self): return self._un(1)
    def sqrt(self): return self._un(2)

    def _reduce(self, op, per_column=False):
        out = np.zeros(self.cols if per_column else 1, dtype=np.float64)
        call("ibx_reduce", context(), op, self.h, int(per_column), ptr(out))
        return out if per_column else out[0]

    def sum(self, per_column=False): return self._reduce(0, per_column)
    def max(self): return F32(self._reduce(1))
    def min(self): return F32(self._reduce(2))
    def maxabs(self): return F32(self._reduce(3))
    def norm(self): return F32(np.sqrt(self._reduce(4)))

    def col(self, j):
        """Copy of column j as a vector."""
        out = DeviceArray(self.rows, 1, True)
        call("ibx_array_column", context(), self.h, int(j), out.h)
        return out


def maximum(a, b):
    return a._bin(4, b) if isinstance(a, DeviceArray) else b._bin(4, a, True)


def minimum(a, b):
    return a._bin(5, b) if isinstance(a, DeviceArray) else b._bin(5, a, True)


def dot(a, b):
    out = C.c_double()
    call("ibx_dot", context(), a.h, b.h, C.byref(out))
    return out.value


# --------------------------------------------------------------------------------- accumulators
class Accumulator:
    """``Accumulator`` (``src/accumulator.jl:12-130``) backed by device CSR tables."""

    def __init__(self, ptr_=None, idx=None, w=None, _handle=None, _owned=True):
        self._h = C.c_void_p()
        self._owned = _owned
        if _handle is not None:
            self._h = _handle
        else:
            p = np.ascontiguousarray(ptr_, dtype=I32)
            i = np.ascontiguousarray(idx, dtype=I32)
            ww = None if w is None else np.ascontiguousarray(w, dtype=F32)
            call("ibx_accum_create", len(p) - 1, ptr(p), ptr(i), ptr(ww), C.byref(self._h))
        n, nnz, wt = C.c_int64(), C.c_int64(), C.c_int()
        call("ibx_accum_info", self._h, C.byref(n), C.byref(nnz), C.byref(wt))
        self.n_output, self.nnz, self.weighted = n.value, nnz.value, bool(wt.value)
        self._uploaded = False

    @classmethod
    def from_lists(cls, inds, weights=None):
        """``Accumulator(inds, weights)`` (``src/accumulator.jl:39-65``), 0-based stencils."""
        lens = np.fromiter((len(s) for s in inds), dtype=np.int64, count=len(inds))
        p = np.concatenate([[0], np.cumsum(lens)])
        idx = np.concatenate([np.asarray(s, dtype=I32) for s in inds]) if len(inds) else np.zeros(0, I32)
        w = None if weights is None else np.concatenate([np.asarray(s, dtype=F32) for s in weights])
        return cls(p, idx, w)

    def __del__(self):
        try:
            if self._owned:
                _lib.lib.ibx_accum_free(self._h)
        except Exception:
            pass

    def tables(self):
        p = np.zeros(self.n_output + 1, dtype=I32)
        i = np.zeros(self.nnz, dtype=I32)
        w = np.zeros(self.nnz, dtype=F32) if self.weighted else None
        call("ibx_accum_tables", self._h, ptr(p), ptr(i), ptr(w))
        return p, i, w

    _F = {None: 0, "identity": 0, "abs": 1, "abs2": 2, "square": 2, "sign": 3}
    _OP = {None: 0, "+": 0, "sum": 0, "max": 1, "min": 2, "*": 3, "prod": 3}

    def __call__(self, v, delta=False, f=None, op=None):
        """``acc(v; Δ, f, op)`` (``src/accumulator.jl:78-130``); v: DeviceArray or host array.  ``f`` / ``op`` by name
        (``"abs"``, ``"abs2"``, ``"sign"``; ``"+"``, ``"max"``, ``"min"``, ``"*"``): closures do not cross the C ABI."""
        host = not isinstance(v, DeviceArray)
        dv = DeviceArray.from_host(v) if host else v
        if not self._uploaded:
            call("ibx_accum_upload", context(), self._h)
            self._uploaded = True
        out = DeviceArray(self.n_output, dv.cols, dv.vector)
        if f is None and op is None:
            call("ibx_accumulate", context(), self._h, dv.h, int(delta), out.h)
        else:
            call("ibx_accumulate_ex", context(), self._h, dv.h, int(delta), self._F[f], self._OP[op], out.h)
        return out.to_host() if host else out


def Interpolator(X, Xc, bias=None, linear=True, k=0):
    """``Interpolator(X, Xc; bias, linear, k)`` (``src/nninterp.jl:85-138``) -> ``Accumulator``; point index first."""
    X = np.ascontiguousarray(X, dtype=F32)
    Xc = np.ascontiguousarray(Xc, dtype=F32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=F32)
    out = C.c_void_p()
    call("ibx_interpolator_build", X.shape[1], X.shape[0], ptr(X), Xc.shape[0], ptr(Xc), ptr(b), int(linear), int(k),
         C.byref(out))
    return Accumulator(_handle=out)


# --------------------------------------------------------------------------------- domain pieces
class Partition:
    """``Partition`` (``src/ImmersedBoundary.jl:383-392``): host copies of the tables + device operators."""

    def __init__(self, dom, p):
        self.dom, self.p, self.id = dom, p, p + 1
        nd = dom.ndims
        n_dom, n_img, start = C.c_int64(), C.c_int64(), C.c_int64()
        nf = (C.c_int64 * nd)()
        call("ibx_partition_info", dom._h, p, C.byref(n_dom), C.byref(n_img), C.byref(start), nf)
        self.n_domain, self.n_image, self.image_start = n_dom.value, n_img.value, start.value
        self.nfaces = [nf[d] for d in range(nd)]
        self._tables = None

    @property
    def ndims(self):
        return self.dom.ndims

    def tables(self):
        """Host copies: domain, image, image_in_domain, owners/neighbors per dim, left/right CSR lists."""
        if self._tables is None:
            nd = self.ndims
            domain = np.zeros(self.n_domain, dtype=I32)
            iid = np.zeros(self.n_image, dtype=I32)
            call("ibx_partition_tables", self.dom._h, self.p, ptr(domain), ptr(iid))
            t = dict(domain=domain, image_in_domain=iid,
                     image=np.arange(self.image_start, self.image_start + self.n_image, dtype=I32), faces={}, lists={})
            for dim in range(nd):
                o = np.zeros(self.nfaces[dim], dtype=I32)
                n = np.zeros(self.nfaces[dim], dtype=I32)
                call("ibx_partition_faces", self.dom._h, self.p, dim, ptr(o), ptr(n))
                t["faces"][dim] = (o, n)
                for side in (0, 1):
                    pp = np.zeros(self.n_domain + 1, dtype=I32)
                    call("ibx_partition_face_lists", self.dom._h, self.p, dim, side, ptr(pp), None)
                    ii = np.zeros(pp[-1], dtype=I32)
                    call("ibx_partition_face_lists", self.dom._h, self.p, dim, side, ptr(pp), ptr(ii))
                    t["lists"][(dim, bool(side))] = (pp, ii)
            self._tables = t
        return self._tables

    @property
    def domain(self):
        return self.tables()["domain"]

    @property
    def image(self):
        return self.tables()["image"]

    @property
    def image_in_domain(self):
        return self.tables()["image_in_domain"]

    @property
    def spacing(self):
        out = DeviceArray(self.n_domain, self.ndims, False)
        call("ibx_partition_spacing", context(), self.dom._h, self.p, out.h)
        return out

    @property
    def centers(self):
        out = DeviceArray(self.n_domain, self.ndims, False)
        call("ibx_partition_centers", context(), self.dom._h, self.p, out.h)
        return out


class Boundary:
    """``Boundary`` (``src/ImmersedBoundary.jl:406-414``), one chunk of <= max_partition_size ghosts."""

    def __init__(self, dom, b, part):
        self.dom, self.b, self.part = dom, b, part
        nd = dom.ndims
        g, m, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        call("ibx_boundary_info", dom._h, b, part, C.byref(g), C.byref(m), C.byref(nnz))
        G, M, NZ = g.value, m.value, nnz.value
        self.ghost_indices = np.zeros(G, dtype=I32)
        self.projections = np.zeros((G, nd), dtype=F32)
        self.normals_host = np.zeros((G, nd), dtype=F32)
        self.image_distances = np.zeros(G, dtype=F32)
        self.ghost_distances = np.zeros(G, dtype=F32)
        self.image_domain = np.zeros(M, dtype=I32)
        self.interp_ptr = np.zeros(G + 1, dtype=I32)
        self.interp_idx = np.zeros(NZ, dtype=I32)
        self.interp_w = np.zeros(NZ, dtype=F32)
        call("ibx_boundary_tables", dom._h, b, part, ptr(self.ghost_indices), ptr(self.projections),
             ptr(self.normals_host), ptr(self.image_distances), ptr(self.ghost_distances), ptr(self.image_domain),
             ptr(self.interp_ptr), ptr(self.interp_idx), ptr(self.interp_w))

    @property
    def nghost(self):
        return len(self.ghost_indices)

    @property
    def n_tied(self):
        """Ghosts whose k-th / (k+1)-th donor candidates are exactly equidistant (``ibx_boundary_tie_count``)."""
        t = C.c_int64()
        call("ibx_boundary_tie_count", self.dom._h, self.b, self.part, C.byref(t))
        return t.value

    @property
    def normals(self):
        """Device copy (nghost x nd) for BC closures."""
        out = DeviceArray(self.nghost, self.dom.ndims, False)
        call("ibx_bc_normals", context(), self.dom._h, self.b, self.part, out.h)
        return out


class Surface:
    """``Surface`` (``src/ImmersedBoundary.jl:335-376``)."""

    def __init__(self, dom, s):
        self.dom, self.s = dom, s
        nd = dom.ndims
        n, nz, nzo = C.c_int64(), C.c_int64(), C.c_int64()
        call("ibx_surface_info", dom._h, s, C.byref(n), C.byref(nz), C.byref(nzo))
        N = n.value
        self.points = np.zeros((N, nd), dtype=F32)
        self.offsets = np.zeros(N, dtype=F32)
        self.normals = np.zeros((N, nd), dtype=F32)
        self.areas = np.zeros(N, dtype=F32)
        self.ptr, self.idx, self.w = np.zeros(N + 1, I32), np.zeros(nz.value, I32), np.zeros(nz.value, F32)
        self.optr, self.oidx, self.ow = np.zeros(N + 1, I32), np.zeros(nzo.value, I32), np.zeros(nzo.value, F32)
        call("ibx_surface_tables", dom._h, s, ptr(self.points), ptr(self.offsets), ptr(self.normals), ptr(self.areas),
             ptr(self.ptr), ptr(self.idx), ptr(self.w), ptr(self.optr), ptr(self.oidx), ptr(self.ow))

    def _values(self, u, offset):
        host = not isinstance(u, DeviceArray)
        du = DeviceArray.from_host(u) if host else u
        out = DeviceArray(len(self.offsets), du.cols, du.vector)
        call("ibx_surface_values", context(), self.dom._h, self.s, int(offset), du.h, out.h)
        return out.to_host() if host else out

    def __call__(self, u):
        """``(surf::Surface)(u)`` (``:368``)."""
        return self._values(u, 0)

    def at_offset(self, u):
        """``at_offset(surf, u)`` (``:376``)."""
        return self._values(u, 1)


def at_offset(surf, u):
    return surf.at_offset(u)


def surface_integral(surf, u):
    """``surface_integral`` (``src/ImmersedBoundary.jl:351-361``); u: values at the surface points."""
    host = not isinstance(u, DeviceArray)
    du = DeviceArray.from_host(u) if host else u
    out = np.zeros(du.cols, dtype=F32)
    call("ibx_surface_integral", context(), surf.dom._h, surf.s, du.h, ptr(out))
    return out[0] if du.vector else out


class Domain:
    """``Domain(msh; max_partition_size, partition_skirt_depth, ghost_layer_ratio, hypercube_families)``
    (``src/ImmersedBoundary.jl:483-786``).  ``hypercube_families`` is a list of
    ``(name, [(dim0, front), ...])`` with 0-based dims.

    ``build_partitions=False`` skips the per-partition face tables (only the fused, block-structured
    entry points are usable then); ``upload=False`` keeps everything on the host (table inspection
    on a machine without a GPU)."""

    def __init__(self, msh, max_partition_size=100_000, partition_skirt_depth=2, ghost_layer_ratio=F32(1.5),
                 hypercube_families=(), build_partitions=True, build_surfaces=True, upload=True, _handle=None,
                 for_rank=None):
        self.mesh = msh
        self._h = C.c_void_p()
        fams = list(hypercube_families)
        if _handle is not None:
            self._h = _handle
            self._describe(max_partition_size, partition_skirt_depth, ghost_layer_ratio, fams)
            return
        names = (C.c_char_p * max(len(fams), 1))(*[n.encode() for n, _ in fams])
        fptr = np.zeros(len(fams) + 1, dtype=np.int32)
        dims, fronts = [], []
        for i, (_, faces) in enumerate(fams):
            for d, front in faces:
                dims.append(int(d))
                fronts.append(int(bool(front)))
            fptr[i + 1] = len(dims)
        dims = np.asarray(dims if dims else [0], dtype=np.int32)
        fronts = np.asarray(fronts if fronts else [0], dtype=np.int32)
        if for_rank is not None:  # (rank, nranks): ghosts only for that rank's block range, input of .shard()
            call("ibx_domain_build_for_rank", msh._h, float(ghost_layer_ratio), len(fams), names, ptr(fptr), ptr(dims),
                 ptr(fronts), int(for_rank[0]), int(for_rank[1]), C.byref(self._h))
        else:
            call("ibx_domain_build", msh._h, int(max_partition_size), int(partition_skirt_depth), float(ghost_layer_ratio),
                 len(fams), names, ptr(fptr), ptr(dims), ptr(fronts), int(build_partitions), int(build_surfaces),
                 C.byref(self._h))
        self._describe(max_partition_size, partition_skirt_depth, ghost_layer_ratio, fams)
        if upload:
            self.upload()

    def _describe(self, max_partition_size, partition_skirt_depth, ghost_layer_ratio, fams):
        nd, nc, nf, npart, nb, ns = C.c_int(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int()
        call("ibx_domain_info", self._h, C.byref(nd), C.byref(nc), C.byref(nf), C.byref(npart), C.byref(nb), C.byref(ns))
        self.ndims, self.ncells, self.nfaces = nd.value, nc.value, nf.value
        t21, nblk = C.c_int(), C.c_int64()
        call("ibx_domain_flags", self._h, C.byref(t21), C.byref(nblk))
        self.two_to_one, self.nblocks = bool(t21.value), nblk.value
        self.partitions = {p + 1: Partition(self, p) for p in range(npart.value)}
        self.boundaries, self.boundary_index = {}, {}
        for b in range(nb.value):
            nm, nparts = C.c_char_p(), C.c_int()
            call("ibx_boundary_name", self._h, b, C.byref(nm), C.byref(nparts))
            name = nm.value.decode()
            self.boundary_index[name] = b
            self.boundaries[name] = {k + 1: Boundary(self, b, k) for k in range(nparts.value)}
        self.surfaces = {}
        for s in range(ns.value):
            nm = C.c_char_p()
            call("ibx_surface_name", self._h, s, C.byref(nm))
            self.surfaces[nm.value.decode()] = Surface(self, s)
        self.reconstruction_kwargs = dict(max_partition_size=max_partition_size,
                                          partition_skirt_depth=partition_skirt_depth,
                                          ghost_layer_ratio=ghost_layer_ratio,
                                          hypercube_families=[(n, list(f)) for n, f in fams])
        self.uploaded = False
        self.shard_info = None

    # -- multi-GPU: rank-local shard
    def shard(self, rank, nranks, all_gather_object=None):
        """Rank-local domain: the rank's contiguous block range + the skirt / image-donor cells it reads.

        ``all_gather_object(obj) -> [obj_rank0, ...]`` moves the halo request lists between ranks (e.g.
        ``torch.distributed.all_gather_object``); with ``nranks == 1`` it is not needed."""
        out = C.c_void_p()
        call("ibx_domain_shard", self._h, int(rank), int(nranks), C.byref(out))
        loc = Domain(self.mesh, _handle=out, **{k: v for k, v in self.reconstruction_kwargs.items()})
        no, nh, st = C.c_int64(), C.c_int64(), C.c_int64()
        call("ibx_shard_info", loc._h, C.byref(no), C.byref(nh), C.byref(st))
        l2g = np.zeros(no.value + nh.value, dtype=I32)
        call("ibx_shard_tables", loc._h, ptr(l2g))
        sc, rc = np.zeros(nranks, np.int64), np.zeros(nranks, np.int64)
        call("ibx_halo_sizes", loc._h, nranks, ptr(sc), ptr(rc))
        requests = {}
        for peer in range(nranks):
            r = np.zeros(rc[peer], dtype=I32)
            call("ibx_halo_lists", loc._h, peer, None, ptr(r))
            requests[peer] = l2g[r]          # global ids, in this rank's unpack order
        loc.shard_info = dict(rank=rank, nranks=nranks, n_owned=no.value, n_halo=nh.value, owned_start=st.value,
                              local_to_global=l2g, requests=requests)
        if nranks > 1:
            everyone = all_gather_object(requests)
            for peer in range(nranks):
                ids = np.ascontiguousarray(everyone[peer][rank], dtype=I32)
                call("ibx_shard_set_send", loc._h, peer, len(ids), ptr(ids))
            loc.shard_info["coupled_families"] = loc._coupled_families(all_gather_object)
        return loc

    def _coupled_families(self, all_gather_object):
        """Ordered pairs (e, l) of boundary families such that, on some rank, a ghost of family l interpolates from a
        cell owned by ANOTHER rank that is a ghost of family e there.  The reference applies its ``impose_bc!`` calls
        one after the other on one array, so family l must see family e's new values: ``ghost_update_euler`` exchanges
        the halo rows between two such families (within one family the update is Jacobi: all reads before any write)."""
        info = self.shard_info
        l2g, n_owned = info["local_to_global"], info["n_owned"]
        halo_donors, mine = {}, {}
        for name, chunks in self.boundaries.items():
            d = np.concatenate([b.image_domain for b in chunks.values()] or [np.zeros(0, I32)])
            halo_donors[name] = np.unique(l2g[d[d >= n_owned]])
            mine[name] = np.unique(l2g[np.concatenate([b.ghost_indices for b in chunks.values()] or [np.zeros(0, I32)])])
        pairs = set()
        for peer, theirs in enumerate(all_gather_object(halo_donors)):
            if peer == info["rank"]:
                continue
            for l, ids in theirs.items():
                pairs |= {(e, l) for e, g in mine.items() if e != l and len(ids) and np.isin(ids, g, assume_unique=True).any()}
        return set().union(*all_gather_object(sorted(pairs)))

    def halo_exchange(self, a):
        """Post and complete the exchange of the halo rows of device array ``a`` (n_owned + n_halo rows)."""
        call("ibx_halo_begin", context(), self._h, a.h)
        call("ibx_halo_end", context(), self._h, a.h)

    def halo_begin(self, a):
        """Post the exchange only.  ``residual_euler`` on the same array completes it after converting the owned rows
        (the exchange then overlaps that kernel); any other consumer must call ``halo_end`` first."""
        call("ibx_halo_begin", context(), self._h, a.h)

    def halo_end(self, a):
        call("ibx_halo_end", context(), self._h, a.h)

    def send_lists(self):
        out = {}
        n = self.shard_info["nranks"]
        sc, rc = np.zeros(n, np.int64), np.zeros(n, np.int64)
        call("ibx_halo_sizes", self._h, n, ptr(sc), ptr(rc))
        for peer in range(n):
            s_, r_ = np.zeros(sc[peer], dtype=I32), np.zeros(rc[peer], dtype=I32)
            call("ibx_halo_lists", self._h, peer, ptr(s_), ptr(r_))
            out[peer] = (s_, r_)
        return out

    def upload(self):
        if not self.uploaded:
            call("ibx_domain_upload", context(), self._h)
            self.uploaded = True

    def __del__(self):
        try:
            _lib.lib.ibx_domain_free(self._h)
        except Exception:
            pass

    def __len__(self):
        """``Base.length(::Domain)`` (``src/ImmersedBoundary.jl:871``)."""
        return self.ncells

    # -- host copies of global tables (tests / inspection)
    def faces(self):
        out = np.zeros((self.nfaces, 3), dtype=I32)
        call("ibx_domain_faces", self._h, ptr(out))
        return out

    def cells(self):
        c = np.zeros((self.ncells, self.ndims), dtype=F32)
        w = np.zeros((self.ncells, self.ndims), dtype=F32)
        call("ibx_domain_cells", self._h, ptr(c), ptr(w))
        return c, w

    def block_faces(self):
        out = np.zeros((self.nblocks, 2 * self.ndims, 7), dtype=I32)
        call("ibx_domain_block_faces", self._h, ptr(out))
        return out

    # -- partition runtime (src/ImmersedBoundary.jl:820-864)
    def __call__(self, f, *args, **kwargs):
        """``dom(f, args...)``: for every partition gather ``a[part.domain, :]`` into device arrays, call
        ``f(part, *dargs, **kwargs)``, scatter the image rows of every arg back.  ``args`` are host float32
        arrays (mutated in place, like the reference) or global ``DeviceArray``s.  Results are returned in
        ascending partition id.  One CUDA stream replaces the reference's ``n_threads`` tasks."""
        self.upload()
        kwargs.pop("n_threads", None)
        gl = [a if isinstance(a, DeviceArray) else DeviceArray.from_host(a) for a in args]
        results = []
        for pid in sorted(self.partitions):
            part = self.partitions[pid]
            dargs = []
            for g in gl:
                da = DeviceArray(part.n_domain, g.cols, g.vector)
                call("ibx_gather_domain", context(), self._h, part.p, g.h, da.h)
                dargs.append(da)
            results.append(f(part, *dargs, **kwargs))
            for g, da in zip(gl, dargs):
                call("ibx_scatter_image", context(), self._h, part.p, da.h, g.h)
        for a, g in zip(args, gl):
            if not isinstance(a, DeviceArray):
                a[...] = g.to_host().reshape(a.shape)
        return results


def volume_integral(dom, A):
    """``volume_integral`` (``src/ImmersedBoundary.jl:1415-1431``)."""
    dom.upload()
    dA = A if isinstance(A, DeviceArray) else DeviceArray.from_host(A)
    out = np.zeros(dA.cols, dtype=F32)
    call("ibx_volume_integral", context(), dom._h, dA.h, ptr(out))
    return out[0] if dA.vector else out


# --------------------------------------------------------------------------------- grid operators
def _op(name, part, dim, u, rows, cols=None, vector=None):
    out = DeviceArray(rows, u.cols if cols is None else cols, u.vector if vector is None else vector, f64=u.f64)
    call(name, context(), part.dom._h, part.p, int(dim), u.h, out.h)
    return out


def at_owners(part, u, dim):
    """``at_owners`` (``src/ImmersedBoundary.jl:879-881``)."""
    return _op("ibx_at_owners", part, dim, u, part.nfaces[dim])


def at_neighbors(part, u, dim):
    """``at_neighbors`` (``:889-891``)."""
    return _op("ibx_at_neighbors", part, dim, u, part.nfaces[dim])


def at_faces(part, u, dim):
    """``at_faces`` (``:899-910``)."""
    return _op("ibx_at_faces", part, dim, u, part.nfaces[dim])


def green_gauss(part, uf, dim):
    """``green_gauss`` (``:918-926``)."""
    return _op("ibx_green_gauss", part, dim, uf, part.n_domain)


def unsigned_green_gauss(part, uf, dim):
    """``unsigned_green_gauss`` (``:934-942``)."""
    return _op("ibx_unsigned_green_gauss", part, dim, uf, part.n_domain)


def divergent(part, uf):
    """``divergent`` (``:950-956``)."""
    s = green_gauss(part, uf[0], 0)
    for dim in range(1, part.ndims):
        s += green_gauss(part, uf[dim], dim)
    return s


def cell_gradient(part, u, dim=None):
    """``cell_gradient`` (``:965-987``)."""
    if dim is None:
        return tuple(cell_gradient(part, u, d) for d in range(part.ndims))
    return _op("ibx_cell_gradient", part, dim, u, part.n_domain)


def _dist(name, part, dim):
    out = DeviceArray(part.nfaces[dim], 1, True)
    call(name, context(), part.dom._h, part.p, int(dim), out.h)
    return out


def face_distance(part, dim):
    """``face_distance`` (``:995-1002``)."""
    return _dist("ibx_face_distance", part, dim)


def owner_distance(part, dim):
    """``owner_distance`` (``:1010-1016``)."""
    return _dist("ibx_owner_distance", part, dim)


def neighbor_distance(part, dim):
    """``neighbor_distance`` (``:1024-1030``)."""
    return _dist("ibx_neighbor_distance", part, dim)


def face_gradient(part, u, dim, grad_u=None):
    """``face_gradient`` (``:1039-1069``)."""
    if grad_u is None:
        return _op("ibx_face_gradient", part, dim, u, part.nfaces[dim])
    return tuple(face_gradient(part, u, dim) if i == dim else at_faces(part, grad_u[i], dim)
                 for i in range(part.ndims))


def JST_sensor(part, p, dim=None):
    """``JST_sensor(part, p, dim)`` (``:1077-1097``); ``dim=None`` is the reference's ``dim = 0``."""
    return _op("ibx_jst_sensor", part, -1 if dim is None else dim, p, part.n_domain)


def MUSCL(part, u, du, dim, D=None, high_order=False):
    """``MUSCL`` (``:1113-1157``) -> (uL, uR)."""
    nf = part.nfaces[dim]
    uL, uR = DeviceArray(nf, u.cols, u.vector), DeviceArray(nf, u.cols, u.vector)
    call("ibx_muscl", context(), part.dom._h, part.p, int(dim), u.h, du.h, 0 if D is None else D.h, int(high_order),
         uL.h, uR.h)
    return uL, uR


# --------------------------------------------------------------------------------- IB ghost update
def impose_bc(f, dom, bname, *args, **kwargs):
    """``impose_bc!(f, dom, bname, args...)`` (``src/ImmersedBoundary.jl:1197-1247``).

    ``f(bdry, *image_values)`` returns a scalar, a ``DeviceArray`` or a tuple of them; ``args`` are host
    arrays (mutated in place) or global ``DeviceArray``s.  All boundary partitions read before any writes."""
    dom.upload()
    kwargs.pop("n_threads", None)
    gl = [a if isinstance(a, DeviceArray) else DeviceArray.from_host(a) for a in args]
    b = dom.boundary_index[bname]
    pending = []
    for k in sorted(dom.boundaries[bname]):
        bdry = dom.boundaries[bname][k]
        iargs = []
        for g in gl:
            ia = DeviceArray(bdry.nghost, g.cols, g.vector)
            call("ibx_bc_image_values", context(), dom._h, b, bdry.part, g.h, ia.h)
            iargs.append(ia)
        r = f(bdry, *iargs, **kwargs)
        if not isinstance(r, tuple):
            r = (r,)
        pending.append((bdry, iargs, r))
    for bdry, iargs, r in pending:
        for g, ba, ia in zip(gl, r, iargs):
            if isinstance(ba, DeviceArray):
                call("ibx_bc_blend", context(), dom._h, b, bdry.part, g.h, ia.h, ba.h)
            elif np.ndim(ba) == 0:
                call("ibx_bc_blend_scalar", context(), dom._h, b, bdry.part, g.h, ia.h, float(ba))
            else:
                dba = DeviceArray.from_host(np.asarray(ba, dtype=F32))
                call("ibx_bc_blend", context(), dom._h, b, bdry.part, g.h, ia.h, dba.h)
    for a, g in zip(args, gl):
        if not isinstance(a, DeviceArray):
            a[...] = g.to_host().reshape(a.shape)


# --------------------------------------------------------------------------------- multigrid builder
def multigrid(dom, max_levels=0, factor=2):
    """``multigrid(dom)`` (``src/ImmersedBoundary.jl:1355-1407``).

    Returns ``(coarse_doms, prolongators, coarseners)`` -- the order the reference *code* returns
    (``:1406``); its docstring and ``test/rae2822.jl:36`` name them in the other order."""
    msh = dom.mesh
    mdepth = int(np.floor(np.log2(msh.block_size)))
    max_levels = mdepth if max_levels == 0 else max_levels
    coarse_doms, coarseners, prolongators = [], [], []
    Xold = dom.cells()[0]
    bsize = msh.block_size
    for _ in range(max_levels):
        bsize //= factor
        cdom = Domain(msh.coarsened(bsize), upload=dom.uploaded, **dom.reconstruction_kwargs)
        X = cdom.cells()[0]
        coarseners.append(Interpolator(Xold, X, linear=False))
        prolongators.append(Interpolator(X, Xold, linear=False))
        coarse_doms.append(cdom)
        Xold = X
    return coarse_doms, prolongators, coarseners
