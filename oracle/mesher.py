"""Oracle (test infrastructure) restatement of ``src/mesher.jl`` (module BlockMesher).

Layout differences from the reference: point arrays are (npoints, nd) (the
reference is (nd, npoints)); simplices are (nsimplices, nverts), 0-based.
Element types follow the inputs exactly as in Julia (a ``Float64`` literal matrix
stays float64, ``.dat``/STL files are Float32), because several refinement
comparisons depend on which side of a rounding a value falls.
"""
import struct

import numpy as np

from .nninterp import KDTree, pinv, _d2

F32 = np.float32


def norm(v):
    """``LinearAlgebra.norm`` for short vectors: float64 accumulation, result in eltype."""
    v = np.asarray(v)
    dt = v.dtype if v.dtype.kind == "f" else np.float64
    return dt.type(np.sqrt(np.sum(v.astype(np.float64) ** 2)))


def normrows(a):
    a = np.asarray(a)
    dt = a.dtype if a.dtype.kind == "f" else np.float64
    return np.sqrt(np.sum(a.astype(np.float64) ** 2, axis=-1)).astype(dt)


# ---------------------------------------------------------------- distance functions
class Box:
    """``Box``, ``src/mesher.jl:27-46``."""

    def __init__(self, origin, widths):
        self.origin = np.asarray(origin, dtype=np.float64) if not isinstance(origin, np.ndarray) else origin
        self.widths = np.asarray(widths, dtype=np.float64) if not isinstance(widths, np.ndarray) else widths

    def __call__(self, pt):
        d = pt - self.origin
        out = (d > self.widths) | (pt < self.origin)
        return norm(np.minimum(np.abs(d), np.abs(d - self.widths)) * out)


class Ball:
    """``Ball``, ``src/mesher.jl:58-76`` (returns Float64 because of the ``0.0`` literal)."""

    def __init__(self, center, radius):
        self.center = np.asarray(center, dtype=np.float64) if not isinstance(center, np.ndarray) else center
        self.radius = radius

    def __call__(self, pt):
        return np.float64(max(0.0, norm(self.center - pt) - self.radius))


class Line:
    """``Line``, ``src/mesher.jl:94-122`` (``m \\ v`` is the least-squares scalar)."""

    def __init__(self, p1, p2):
        self.p1 = np.asarray(p1, dtype=np.float64) if not isinstance(p1, np.ndarray) else p1
        self.p2 = np.asarray(p2, dtype=np.float64) if not isinstance(p2, np.ndarray) else p2
        self.m = self.p2 - self.p1

    def __call__(self, pt):
        v = pt - self.p1
        xi = np.dot(self.m / np.dot(self.m, self.m), v)  # `m \\ v` == pinv(m) * v
        if xi < 0.0:
            return norm(pt - self.p1)
        if xi > 1.0:
            return norm(pt - self.p2)
        return norm(pt - (self.p1 + self.m * xi))


# ---------------------------------------------------------------- stereolitography
class Stereolitography:
    """``Stereolitography``, ``src/mesher.jl:238-296``.  points (np, nd), simplices (ns, nv)."""

    def __init__(self, points, simplices=None, closed=True):
        if isinstance(points, str):
            p, s = _read_surface(points)
            self.points, self.simplices = p, s
            return
        points = np.asarray(points)
        if simplices is None:
            n = points.shape[0]
            inds = np.arange(n)
            if closed:
                simplices = np.stack([inds, np.roll(inds, -1)], axis=1)
            else:
                simplices = np.stack([inds[:-1], inds[1:]], axis=1)
        self.points = points
        self.simplices = np.asarray(simplices, dtype=np.int64)

    @property
    def nd(self):
        return self.points.shape[1]


def _read_surface(fname):
    """File constructor, ``src/mesher.jl:279-296`` and ``STLReader`` ``:124-227``."""
    if fname[-4:] in (".dat", ".DAT"):
        pts = np.loadtxt(fname, dtype=np.float64).astype(F32)
        n = pts.shape[0]
        inds = np.arange(n)
        return pts, np.stack([inds, np.roll(inds, -1)], axis=1)
    with open(fname, "rb") as fh:
        head = fh.read(5)
    if head == b"solid":
        verts, faces, face = [], [], []
        with open(fname, "r") as fh:
            for line in fh:
                line = line.strip()
                if line.startswith("vertex"):
                    c = line.split()
                    verts.append([F32(c[1]), F32(c[2]), F32(c[3])])
                    face.append(len(verts) - 1)
                elif line.startswith("facet normal"):
                    face = []
                elif line.startswith("endloop"):
                    faces.append(face)
        return np.asarray(verts, dtype=F32), np.asarray(faces, dtype=np.int64)
    with open(fname, "rb") as fh:
        raw = fh.read()
    ntri = struct.unpack_from("<I", raw, 80)[0]
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
    pts = rec["v"].reshape(-1, 3).astype(F32)
    return pts, np.arange(3 * ntri, dtype=np.int64).reshape(ntri, 3)


def write_stl_binary(fname, points, simplices):
    """Test helper (no reference counterpart): write a binary STL."""
    tri = np.asarray(points, dtype=F32)[np.asarray(simplices)]
    rec = np.zeros(len(tri), dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    rec["v"] = tri
    with open(fname, "wb") as fh:
        fh.write(b"\0" * 80)
        fh.write(struct.pack("<I", len(tri)))
        fh.write(rec.tobytes())


def merge_points(*stls, tolerance=1e-7, clean_degenerate=True):
    """``merge_points``, ``src/mesher.jl:351-407``."""
    if not isinstance(tolerance, np.floating):
        tolerance = np.float64(tolerance)  # a Julia Float64 literal promotes Float32 points (NumPy's would not)
    tag2ind = {}
    new_points = []
    new_simplices = []
    for stl in stls:
        tags = np.rint(stl.points / tolerance).astype(np.int64)
        new_idx = np.empty(len(tags), dtype=np.int64)
        for i, t in enumerate(map(tuple, tags.tolist())):
            j = tag2ind.get(t)
            if j is None:
                j = len(new_points)
                tag2ind[t] = j
                new_points.append(stl.points[i])
            new_idx[i] = j
        new_simplices.append(new_idx[stl.simplices])
    pts = np.stack(new_points, axis=0)
    simp = np.concatenate(new_simplices, axis=0)
    if clean_degenerate:
        srt = np.sort(simp, axis=1)
        ok = np.all(srt[:, 1:] != srt[:, :-1], axis=1)
        simp = simp[ok]
    return Stereolitography(pts, simp)


def cat(*stls):
    """``Base.cat(::Stereolitography...)``, ``src/mesher.jl:415-431``."""
    pts = np.concatenate([s.points for s in stls], axis=0)
    off = np.cumsum([0] + [s.points.shape[0] for s in stls[:-1]])
    simp = np.concatenate([s.simplices + o for s, o in zip(stls, off)], axis=0)
    return Stereolitography(pts, simp)


def _refine_simplex(simplex, h, growth_ratio, refinement_regions, out):
    """``refine_to_length!``, ``src/mesher.jl:438-495``; simplex (nv, nd), appended depth-first."""
    stack = [simplex]
    gm1 = np.float64(growth_ratio) - 1.0 if not isinstance(growth_ratio, np.floating) else np.float64(growth_ratio) - 1.0
    while stack:
        s = stack.pop()
        nv = s.shape[0]
        max_violation = 0.0
        index = -1
        for i in range(nv):
            inext = 0 if i == nv - 1 else i + 1
            p1, p2 = s[i], s[inext]
            phalf = (p1 + p2) / 2
            L = norm(p2 - p1)
            hloc = h
            for df, href in refinement_regions:
                cand = max((df(phalf) - L) * gm1, href)
                hloc = min(hloc, cand)
            violation = L - hloc
            if max_violation < violation:
                max_violation = violation
                index = i
        if index < 0:
            out.append(s)
            continue
        inext = 0 if index == nv - 1 else index + 1
        pnew = (s[index] + s[inext]) / 2
        new_s = s.copy()
        s = s.copy()
        s[inext] = pnew
        new_s[index] = pnew
        stack.append(new_s)  # processed after `s` -> depth-first [first; second] order
        stack.append(s)


def refine_to_length(stl, h, tolerance=1e-7, growth_ratio=1.1, refinement_regions=()):
    """``refine_to_length``, ``src/mesher.jl:503-528``."""
    out = []
    for simp in stl.simplices:
        _refine_simplex(stl.points[simp].copy(), h, growth_ratio, list(refinement_regions), out)
    nv = out[0].shape[0]
    pts = np.concatenate(out, axis=0)
    simplices = np.arange(pts.shape[0], dtype=np.int64).reshape(-1, nv)
    return merge_points(Stereolitography(pts, simplices), tolerance=tolerance)


_SEG_EPS = F32(1e-14)


def proj2simplex_batch(simp, pt):
    """``proj2simplex``, ``src/mesher.jl:544-596``, batched: simp (B, nv, nd), pt (B, nd)."""
    nv = simp.shape[1]
    eps = _SEG_EPS
    if nv == 1:
        return simp[:, 0, :].astype(np.result_type(simp.dtype, pt.dtype))
    if nv == 2:
        p0, p1 = simp[:, 0, :], simp[:, 1, :]
        u = p1 - p0
        xi = np.sum((pt - p0) * u, axis=1) / (np.sum(u * u, axis=1) + eps)
        res = p0 + u * xi[:, None]
        dt = res.dtype
        res = np.where((xi < -eps)[:, None], p0.astype(dt), res)
        res = np.where((xi > 1.0 + eps)[:, None], p1.astype(dt), res)
        return res
    p0 = simp[:, 0, :]
    M = np.swapaxes(simp[:, 1:, :] - p0[:, None, :], 1, 2)  # (B, nd, nv-1)
    rhs = pt - p0
    dt = np.result_type(M.dtype, rhs.dtype)
    nd = simp.shape[2]
    if nv == 3:
        # Canonical rule: LinearAlgebra.pinv goes through LAPACK's SVD, which is neither vendored
        # nor bit-reproducible.  For a full-rank nd x 2 matrix pinv(M) = (M^T M)^-1 M^T, evaluated in Float64 from the
        # rounded entries of M (Gram sums in dimension order, closed-form 2 x 2 inverse) and rounded once; products and
        # sums below are then taken one by one in the promoted type, like Julia's generic matvec.
        M64 = M.astype(np.float64)
        m0, m1 = M64[:, :, 0], M64[:, :, 1]
        ga, gb, gc = m0[:, 0] * m0[:, 0], m0[:, 0] * m1[:, 0], m1[:, 0] * m1[:, 0]
        for d in range(1, nd):
            ga, gb, gc = ga + m0[:, d] * m0[:, d], gb + m0[:, d] * m1[:, d], gc + m1[:, d] * m1[:, d]
        det = ga * gc - gb * gb
        ok = det > 1e-10 * (ga * gc)
        with np.errstate(all="ignore"):
            P0 = (gc[:, None] * m0 - gb[:, None] * m1) / det[:, None]
            P1 = (ga[:, None] * m1 - gb[:, None] * m0) / det[:, None]
        Pm = np.stack([P0, P1], axis=1)                       # (B, 2, nd)
        if not ok.all():
            Pm[~ok] = pinv(M[~ok]).astype(np.float64)
        Pm = Pm.astype(M.dtype).astype(dt)
        r_, Mt = rhs.astype(dt), M.astype(dt)
        xi = np.zeros((len(simp), 2), dtype=dt)
        for j in range(2):
            acc = Pm[:, j, 0] * r_[:, 0]
            for d in range(1, nd):
                acc = acc + Pm[:, j, d] * r_[:, d]
            xi[:, j] = acc
        res = p0.astype(dt) + (Mt[:, :, 0] * xi[:, 0:1] + Mt[:, :, 1] * xi[:, 1:2])
    else:
        xi = np.einsum("bij,bj->bi", pinv(M).astype(dt), rhs.astype(dt))
        res = (p0.astype(dt) + np.einsum("bij,bj->bi", M.astype(dt), xi))
    inside = ~(np.any(xi < -eps, axis=1) | (np.sum(xi, axis=1) > 1.0 + eps))
    out_idx = np.flatnonzero(~inside)
    if out_idx.size:
        best = np.zeros((out_idx.size, simp.shape[2]), dtype=dt)
        bd = np.full(out_idx.size, np.inf, dtype=np.float32).astype(dt)
        for drop in range(nv):  # simplex_faces: all vertices but `drop` (mesher.jl:533-539)
            keep = [j for j in range(nv) if j != drop]
            pf = proj2simplex_batch(simp[out_idx][:, keep, :], pt[out_idx]).astype(dt)
            d = normrows(pf - pt[out_idx])
            better = d < bd
            bd = np.where(better, d, bd)
            best = np.where(better[:, None], pf, best)
        res[out_idx] = best
    return res


def simplex_normals(stl, normalize=False):
    """``_simplex_normal``, ``src/mesher.jl:601-628`` for all simplices."""
    s = stl.points[stl.simplices]
    if stl.nd == 2:
        v = s[:, 1, :] - s[:, 0, :]
        n = np.stack([v[:, 1], -v[:, 0]], axis=1)
        if normalize:
            n = n / (normrows(v) + _SEG_EPS)[:, None]
        return n
    p0 = s[:, 0, :]
    n = np.cross(s[:, 1, :] - p0, s[:, 2, :] - p0)
    if normalize:
        n = n / (normrows(n) + _SEG_EPS)[:, None]
    return n


def centers_and_normals(stl):
    """``centers_and_normals``, ``src/mesher.jl:639-660`` -> (centers (ns, nd), normals (ns, nd))."""
    s = stl.points[stl.simplices]
    acc = s[:, 0, :]
    for j in range(1, s.shape[1]):
        acc = acc + s[:, j, :]
    return acc / s.shape[1], simplex_normals(stl, False)


def feature_regions(stl, angle=15.0, radius=np.inf, include_boundaries=False):
    """``feature_regions``, ``src/mesher.jl:670-728``."""
    eps = np.finfo(np.float32).eps
    ang = np.deg2rad(max(angle, 1.0))
    max_cos = np.cos(np.deg2rad(0.05))
    edges = []
    registry = {}
    for i, simp in enumerate(stl.simplices.tolist()):
        for pivot in simp:
            face = tuple(sorted(v for v in simp if v != pivot))
            if face in registry:
                edges.append((registry.pop(face), i))
            else:
                registry[face] = i
    for ind in registry.values():
        edges.append((ind, ind))
    centers, normals = centers_and_normals(stl)
    included = np.zeros(len(stl.simplices), dtype=bool)
    for i, j in edges:
        ni = normals[i] / (norm(normals[i]) + eps)
        nj = normals[j] / (norm(normals[j]) + eps)
        theta = np.arccos(min(float(np.dot(ni, nj)), max_cos))
        d = float(norm(centers[i] - centers[j]))
        if (i == j and include_boundaries) or (d / theta < radius) or (theta > ang):
            included[i] = True
            included[j] = True
    return Stereolitography(stl.points, stl.simplices[included])


class DistanceField:
    """``DistanceField``, ``src/mesher.jl:736-801``."""

    def __init__(self, stl, leaf_size=25, h=0.0):
        if h > 0.0:
            stl = refine_to_length(stl, h)
        self.stl = stl
        self.centers, _ = centers_and_normals(stl)
        self.tree = KDTree(self.centers, leafsize=leaf_size)

    def __call__(self, x):
        _, d = self.tree.nn(np.asarray(x)[None, :])
        return d[0]

    def distances(self, X):
        return self.tree.nn(X)[1]

    def projection(self, X, R):
        """``projection``, ``src/mesher.jl:778-801``, batched over rows of X with radii R."""
        X = np.asarray(X)
        idx, d = self.tree.nn(X)
        dt = np.result_type(self.centers.dtype, X.dtype)
        p = self.centers[idx].astype(dt)
        d = d.astype(dt)
        R = np.broadcast_to(np.asarray(R), (X.shape[0],))
        need = np.flatnonzero(R > d)
        if need.size == 0:
            return p
        cand = self.tree.inrange_many(X[need], R[need].astype(np.float64))
        qid = np.concatenate([np.full(len(c), q, dtype=np.int64) for q, c in zip(need, cand)])
        sid = np.concatenate([np.sort(np.asarray(c, dtype=np.int64)) for c in cand]) if len(cand) else np.zeros(0, np.int64)
        if sid.size == 0:
            return p
        # exact boundary-inclusive filter in the promoted type (NearestNeighbors.inrange)
        cd2 = _d2(self.centers[sid], X[qid])
        Rq = R[qid].astype(cd2.dtype)
        keep = cd2 <= Rq * Rq
        qid, sid = qid[keep], sid[keep]
        simp = self.stl.points[self.stl.simplices[sid]]
        pr = proj2simplex_batch(simp, X[qid]).astype(dt)
        dd = normrows(pr - X[qid].astype(dt))
        # sequential `if _d < d` over ascending simplex index: first strict minimum wins
        order = np.lexsort((sid, dd, qid))
        qs = qid[order]
        first = np.concatenate([[True], qs[1:] != qs[:-1]])
        bi = order[first]
        bq = qid[bi]
        better = dd[bi] < d[bq]
        p[bq[better]] = pr[bi[better]]
        return p


# ---------------------------------------------------------------- octree
def refine_octree(criteria, origin, widths, growth_ratio=1.1):
    """``refine_octree``, ``src/mesher.jl:811-862``; returns (origins (nb, nd), widths (nb, nd)) float32."""
    out_o, out_w = [], []
    gm1 = np.float64(growth_ratio) - 1.0
    stack = [(np.asarray(origin, dtype=F32), np.asarray(widths, dtype=F32), list(criteria))]
    while stack:
        o, w, crit = stack.pop()
        L = w.max()
        R = norm(w) / F32(2)
        center = o + w / F32(2)
        active = []
        for df, h in crit:
            Lmax = max(gm1 * (df(center) - R), h)
            if Lmax < L:
                active.append((df, h))
        if not active:
            out_o.append(o)
            out_w.append(w)
            continue
        wmin = w.min()
        split = np.rint(w / wmin).astype(np.int64) + 1
        new_w = (w / split).astype(F32)
        axes = []
        for d in range(len(o)):
            a = np.float64(o[d])
            b = np.float64(F32(o[d] + w[d]))
            s = int(split[d])
            t = np.arange(s, dtype=np.float64) / s
            axes.append(((1.0 - t) * a + t * b).astype(F32))  # LinRange lerp, last point dropped
        grids = np.meshgrid(*axes, indexing="ij")
        # Iterators.product order: first dimension fastest
        child_o = np.stack([g.ravel(order="F") for g in grids], axis=1)
        for c in child_o[::-1]:  # reversed push -> popped in product order (depth-first)
            stack.append((c.astype(F32), new_w, active))
    return np.stack(out_o, axis=0), np.stack(out_w, axis=0)


def refine_orderly(surfaces, refinement_regions=(), ratio=F32(0.5), growth_ratio=F32(2.0), tolerance=F32(1e-7)):
    """``refine_orderly``, ``src/mesher.jl:878-918``; surfaces: list of (stl, h)."""
    hs = [s[1] for s in surfaces]
    order = np.argsort(np.asarray(hs, dtype=np.float64), kind="stable")
    regions = [(df, h * ratio) for df, h in refinement_regions]
    result = {}
    for i in order:
        stl, h = surfaces[i]
        h = h * ratio
        if isinstance(stl, AnalyticSphere):
            dfield = stl
        else:
            stl = refine_to_length(stl, h, tolerance=tolerance, refinement_regions=regions, growth_ratio=growth_ratio)
            dfield = DistanceField(stl)
        result[int(i)] = dfield
        regions.append((dfield, h))
    return [result[i] for i in range(len(surfaces))]


class Mesh:
    """``Mesh`` struct and constructor, ``src/mesher.jl:926-1046``."""

    def __init__(self, origin, widths, *surfaces, growth_ratio=F32(2.0), tolerance=F32(1e-7), block_size=8,
                 refinement_regions=(), _raw=None):
        if _raw is not None:
            (self.origin, self.widths, self.block_size, self.block_origins, self.block_widths,
             self.distance_fields) = _raw
            return
        block_size = np.int32(block_size)
        self.origin = np.asarray(origin, dtype=F32)
        self.widths = np.asarray(widths, dtype=F32)
        self.block_size = int(block_size)
        dfields = refine_orderly([(stl, h) for _, stl, h in surfaces], refinement_regions=refinement_regions,
                                 growth_ratio=growth_ratio, tolerance=tolerance)
        self.distance_fields = {s[0]: df for s, df in zip(surfaces, dfields)}
        tb = lambda h: h * type(h)(block_size) if isinstance(h, np.floating) else h * float(block_size)
        regions = [(df, tb(h)) for df, h in refinement_regions]
        for name, _, h in surfaces:
            regions.append((self.distance_fields[name], tb(h)))
        self.block_origins, self.block_widths = refine_octree(regions, self.origin, self.widths, growth_ratio)

    @classmethod
    def from_blocks(cls, origin, widths, block_size, block_origins, block_widths, distance_fields):
        """Positional struct constructor used by ``multigrid`` (``src/ImmersedBoundary.jl:1366-1368``)."""
        return cls(None, None, _raw=(np.asarray(origin, F32), np.asarray(widths, F32), int(block_size),
                                     np.asarray(block_origins, F32), np.asarray(block_widths, F32), distance_fields))

    @property
    def nd(self):
        return self.block_origins.shape[1]

    def __len__(self):
        """``Base.length(::Mesh)``, ``src/ImmersedBoundary.jl:47``."""
        return self.block_size ** self.nd * self.block_origins.shape[0]


def get_cells(msh):
    """``get_cells`` with ``margin = 0``, ``src/mesher.jl:1064-1112``.

    Returns (centers (N, nd), widths (N, nd)) float32; cell order is block-major,
    first dimension fastest inside a block.
    """
    bs = msh.block_size
    nd = msh.nd
    ax = (np.arange(bs, dtype=F32) + F32(0.5)) / F32(bs)
    grids = np.meshgrid(*([ax] * nd), indexing="ij")
    inner = np.stack([g.ravel(order="F") for g in grids], axis=1)  # (bs^nd, nd)
    centers = inner[None, :, :] * msh.block_widths[:, None, :] + msh.block_origins[:, None, :]
    widths = np.repeat((msh.block_widths / F32(bs))[:, None, :], bs ** nd, axis=1)
    return centers.reshape(-1, nd).astype(F32), widths.reshape(-1, nd).astype(F32)


class AnalyticSphere:
    """Oracle twin of the product's analytic sphere surface (extension, see include/ibx.h ``ibx_surface``):
    exact unsigned distance / projection in float64 instead of an STL distance field."""

    stl = None

    def __init__(self, center, radius):
        self.center = np.asarray(center, dtype=np.float64)
        self.radius = float(radius)

    def __call__(self, x):
        return np.float64(abs(np.sqrt(np.sum((np.asarray(x, dtype=np.float64) - self.center) ** 2)) - self.radius))

    def distances(self, X):
        X = np.asarray(X, dtype=np.float64)
        return np.abs(np.sqrt(np.sum((X - self.center) ** 2, axis=1)) - self.radius)

    def projection(self, X, R):
        v = np.asarray(X, dtype=np.float64) - self.center
        s = np.sqrt(np.sum(v ** 2, axis=1))
        return self.center + v * (self.radius / s)[:, None]
