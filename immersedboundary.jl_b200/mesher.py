"""Host-side mirror of the reference's ``BlockMesher`` API (``src/mesher.jl``) on top of the C ABI.

Same names and argument meaning as the Julia package: ``Stereolitography``, ``merge_points``,
``refine_to_length``, ``feature_regions``, ``centers_and_normals``, ``Box``/``Ball``/``Line``,
``DistanceField``, ``Mesh``, ``get_cells``.  Points are (npoints, nd) arrays here (Julia:
(nd, npoints)); simplices are 0-based.  All geometry work happens in libibx (C++).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call, ptr

F32 = np.float32


def _is_f32(x):
    return isinstance(x, np.floating) and x.dtype == np.float32


class Stereolitography:
    """``Stereolitography`` (``src/mesher.jl:238-296``): from a file name, or points (+ simplices)."""

    def __init__(self, points=None, simplices=None, closed=True, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
            return
        if isinstance(points, str):
            call("ibx_stl_read", points.encode(), C.byref(self._h))
            return
        points = np.asarray(points)
        f32 = points.dtype == np.float32
        pts = np.ascontiguousarray(points, dtype=np.float64)
        n, nd = pts.shape
        if simplices is None:
            inds = np.arange(n)
            simplices = (np.stack([inds, np.roll(inds, -1)], 1) if closed else np.stack([inds[:-1], inds[1:]], 1))
        simp = np.ascontiguousarray(simplices, dtype=np.int64)
        call("ibx_stl_create", nd, n, ptr(pts), simp.shape[0], ptr(simp), int(f32), C.byref(self._h))

    def __del__(self):
        try:
            _lib.lib.ibx_stl_free(self._h)
        except Exception:
            pass

    def _info(self):
        nd, np_, ns, f = C.c_int(), C.c_int64(), C.c_int64(), C.c_int()
        call("ibx_stl_info", self._h, C.byref(nd), C.byref(np_), C.byref(ns), C.byref(f))
        return nd.value, np_.value, ns.value, bool(f.value)

    def _copy(self):
        nd, n, ns, f = self._info()
        pts = np.zeros((n, nd), dtype=np.float64)
        simp = np.zeros((ns, nd), dtype=np.int64)
        call("ibx_stl_copy", self._h, ptr(pts), ptr(simp))
        return (pts.astype(F32) if f else pts), simp

    @property
    def points(self):
        return self._copy()[0]

    @property
    def simplices(self):
        return self._copy()[1]


def merge_points(*stls, tolerance=1e-7, clean_degenerate=True):
    """``merge_points`` (``src/mesher.jl:351-407``)."""
    arr = (C.c_void_p * len(stls))(*[s._h for s in stls])
    out = C.c_void_p()
    call("ibx_stl_merge_points", len(stls), arr, float(tolerance), int(_is_f32(tolerance)), int(clean_degenerate),
         C.byref(out))
    return Stereolitography(_handle=out)


def feature_regions(stl, angle=15.0, radius=np.inf, include_boundaries=False):
    """``feature_regions`` (``src/mesher.jl:670-728``)."""
    out = C.c_void_p()
    call("ibx_stl_feature_regions", stl._h, float(angle), float(radius), int(include_boundaries), C.byref(out))
    return Stereolitography(_handle=out)


def centers_and_normals(stl):
    """``centers_and_normals`` (``src/mesher.jl:639-660``) -> (centers, normals), (nsimp, nd) each."""
    nd, _, ns, f = stl._info()
    c = np.zeros((ns, nd))
    n = np.zeros((ns, nd))
    call("ibx_stl_centers_normals", stl._h, ptr(c), ptr(n))
    return (c.astype(F32), n.astype(F32)) if f else (c, n)


class Ball:
    """``Ball(center, radius)`` (``src/mesher.jl:58-76``)."""

    def __init__(self, center, radius):
        self.center, self.radius = np.asarray(center, dtype=np.float64), float(radius)

    def _region(self, h):
        r = _lib.Region(kind=0, h_is_f32=int(_is_f32(h)), h=float(h))
        for d, v in enumerate(self.center):
            r.c[d] = v
        r.a[0] = self.radius
        return r


class Box:
    """``Box(origin, widths)`` (``src/mesher.jl:27-46``)."""

    def __init__(self, origin, widths):
        self.origin, self.widths = np.asarray(origin, dtype=np.float64), np.asarray(widths, dtype=np.float64)

    def _region(self, h):
        r = _lib.Region(kind=1, h_is_f32=int(_is_f32(h)), h=float(h))
        for d in range(len(self.origin)):
            r.c[d], r.a[d] = self.origin[d], self.widths[d]
        return r


class Line:
    """``Line(p1, p2)`` (``src/mesher.jl:94-122``)."""

    def __init__(self, p1, p2):
        self.p1, self.p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)

    def _region(self, h):
        r = _lib.Region(kind=2, h_is_f32=int(_is_f32(h)), h=float(h))
        for d in range(len(self.p1)):
            r.c[d], r.a[d] = self.p1[d], self.p2[d]
        return r


class DistanceField:
    """``DistanceField(stl)`` (``src/mesher.jl:736-769``)."""

    def __init__(self, stl=None, _handle=None, _owner=None):
        self._h = C.c_void_p()
        self._owner = _owner  # keeps the mesh alive when the field belongs to one
        self._own = _handle is None
        if _handle is not None:
            self._h = _handle
        else:
            call("ibx_dfield_create", stl._h, C.byref(self._h))

    def __del__(self):
        try:
            if self._own:
                _lib.lib.ibx_dfield_free(self._h)
        except Exception:
            pass

    @property
    def stl(self):
        out = C.c_void_p()
        call("ibx_dfield_stl", self._h, C.byref(out))
        if not out.value:
            return None
        s = Stereolitography(_handle=out)
        s.__class__ = _BorrowedStl
        s._keep = self
        return s

    def __call__(self, x):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=F32)
        out = np.zeros(x.shape[0])
        call("ibx_dfield_distance", self._h, x.shape[0], ptr(x), ptr(out))
        return out if out.size > 1 else out[0]

    def _region(self, h):
        return _lib.Region(kind=3, h_is_f32=int(_is_f32(h)), h=float(h), dfield=self._h.value)


class _BorrowedStl(Stereolitography):
    def __del__(self):
        pass


def refine_to_length(stl, h, tolerance=1e-7, growth_ratio=1.1, refinement_regions=()):
    """``refine_to_length`` (``src/mesher.jl:503-528``)."""
    regs = (_lib.Region * max(len(refinement_regions), 1))(*[r._region(hr) for r, hr in refinement_regions])
    out = C.c_void_p()
    call("ibx_stl_refine_to_length", stl._h, float(h), int(_is_f32(h)), float(tolerance), int(_is_f32(tolerance)),
         float(growth_ratio), len(refinement_regions), regs, C.byref(out))
    return Stereolitography(_handle=out)


class Sphere:
    """Analytic sphere surface (extension for synthetic throughput meshes; see include/ibx.h)."""

    def __init__(self, center, radius):
        self.center, self.radius = np.asarray(center, dtype=np.float64), float(radius)


class Mesh:
    """``Mesh(origin, widths, surfaces...; refinement_regions, growth_ratio, block_size)``
    (``src/mesher.jl:926-1046``).  ``surfaces`` are ``(name, stl, h)`` tuples; ``refinement_regions``
    are ``(distance_function, h)`` pairs.  Sizes given as ``np.float32`` behave like Julia ``1f-2``
    literals, Python floats like ``1e-2``."""

    def __init__(self, origin=None, widths=None, *surfaces, growth_ratio=F32(2.0), tolerance=F32(1e-7), block_size=8,
                 refinement_regions=(), _handle=None, _surfaces=None):
        self._h = C.c_void_p()
        self._keep = (surfaces, refinement_regions)
        if _handle is not None:
            self._h = _handle
        else:
            o = np.ascontiguousarray(origin, dtype=F32)
            w = np.ascontiguousarray(widths, dtype=F32)
            specs = (_lib.SurfaceSpec * max(len(surfaces), 1))()
            for i, (name, stl, h) in enumerate(surfaces):
                specs[i].name = name.encode()
                specs[i].h = float(h)
                specs[i].h_is_f32 = int(_is_f32(h))
                if isinstance(stl, Sphere):
                    specs[i].stl = None
                    for d in range(3):
                        specs[i].sphere_c[d] = stl.center[d]
                    specs[i].sphere_r = stl.radius
                else:
                    specs[i].stl = stl._h
            regs = (_lib.Region * max(len(refinement_regions), 1))(*[r._region(hr) for r, hr in refinement_regions])
            call("ibx_mesh_create", len(o), ptr(o), ptr(w), len(surfaces), specs, len(refinement_regions), regs,
                 float(growth_ratio), float(tolerance), int(_is_f32(tolerance)), int(block_size), C.byref(self._h))
        nd, bs, nb, nc, ns = C.c_int(), C.c_int(), C.c_int64(), C.c_int64(), C.c_int()
        call("ibx_mesh_info", self._h, C.byref(nd), C.byref(bs), C.byref(nb), C.byref(nc), C.byref(ns))
        self.nd, self.block_size, self.nblocks, self.ncells, self.nsurfaces = nd.value, bs.value, nb.value, nc.value, ns.value
        self.block_origins = np.zeros((self.nblocks, self.nd), dtype=F32)
        self.block_widths = np.zeros((self.nblocks, self.nd), dtype=F32)
        call("ibx_mesh_blocks", self._h, ptr(self.block_origins), ptr(self.block_widths))
        self.distance_fields = {}
        for i in range(self.nsurfaces):
            nm, df = C.c_char_p(), C.c_void_p()
            call("ibx_mesh_surface_name", self._h, i, C.byref(nm))
            call("ibx_mesh_surface_dfield", self._h, i, C.byref(df))
            self.distance_fields[nm.value.decode()] = DistanceField(_handle=df, _owner=self)

    def __del__(self):
        try:
            _lib.lib.ibx_mesh_free(self._h)
        except Exception:
            pass

    def __len__(self):
        """``Base.length(::Mesh)`` (``src/ImmersedBoundary.jl:47``)."""
        return self.ncells

    def save(self, path):
        """Write the mesh (root box, block list, surfaces with their refined STLs) to ``path`` (``ibx_mesh_save``): the
        counterpart of serialising the reference's plain-data ``Mesh`` struct (``src/mesher.jl:926-933``)."""
        call("ibx_mesh_save", self._h, str(path).encode())

    @classmethod
    def load(cls, path):
        """Read a mesh written by ``save`` (``ibx_mesh_load``); distance fields are rebuilt."""
        out = C.c_void_p()
        call("ibx_mesh_load", str(path).encode(), C.byref(out))
        return cls(_handle=out)

    def coarsened(self, block_size):
        """The positional constructor ``multigrid`` uses (``src/ImmersedBoundary.jl:1366-1368``)."""
        out = C.c_void_p()
        call("ibx_mesh_from_blocks", self._h, int(block_size), C.byref(out))
        m = Mesh(_handle=out)
        m._keep = self
        return m


def get_cells(msh):
    """``get_cells`` (``src/mesher.jl:1064-1112``) -> (centers, widths), (ncells, nd) float32."""
    c = np.zeros((msh.ncells, msh.nd), dtype=F32)
    w = np.zeros((msh.ncells, msh.nd), dtype=F32)
    call("ibx_mesh_cells", msh._h, ptr(c), ptr(w))
    return c, w
