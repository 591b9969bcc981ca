"""Oracle (test infrastructure) restatement of ``src/ImmersedBoundary.jl``:
topology build, partition runtime, grid operators, IB ghost update, multigrid builder.

All indices are 0-based (reference: 1-based); "no cell" is -1 (reference: 0).
Field arrays are (cells, nv) / (cells,) float32 with the cell index first.

Canonical face order (the reference's is thread-dependent): interior
faces sorted by (owner, neighbour) -- which is what a single-threaded reference run
produces up to ``inrange``'s traversal order -- followed by the hypercube faces in
the reference's own deterministic order (``src/ImmersedBoundary.jl:157-181``).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .accumulator import Accumulator
from .mesher import Mesh, get_cells, centers_and_normals
from .nninterp import KDTree, Interpolator, interpolator_tables, _d2

F32 = np.float32
EPS32 = np.finfo(np.float32).eps


def _sumsq_rows(a):
    acc = a[:, 0] ** 2
    for d in range(1, a.shape[1]):
        acc = acc + a[:, d] ** 2
    return acc


# ------------------------------------------------------------------ faces
def octree2faces(origins, widths, chunk=20000):
    """``octree2faces``, ``src/ImmersedBoundary.jl:63-132`` -> int64 (nf, 3) rows (dim, owner, neigh)."""
    n, nd = origins.shape
    centers = origins + widths / F32(2)
    tree = KDTree(centers)
    radii = np.sqrt(_sumsq_rows(widths)) / F32(2)
    rr = radii * F32(3.1)
    maxs = origins + widths
    out = []
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lists = tree.inrange_many(centers[s:e], rr[s:e].astype(np.float64))
        cnt = np.fromiter((len(l) for l in lists), dtype=np.int64, count=e - s)
        ii = np.repeat(np.arange(s, e, dtype=np.int64), cnt)
        jj = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if cnt.sum() else np.zeros(0, np.int64)
        keep = ii != jj
        ii, jj = ii[keep], jj[keep]
        d2 = _d2(centers[jj], centers[ii])
        keep = d2 <= rr[ii] * rr[ii]
        ii, jj = ii[keep], jj[keep]
        fo = np.maximum(origins[ii], origins[jj])
        fw = np.minimum(maxs[ii], maxs[jj]) - fo
        tol = F32(0.01) * fw.max(axis=1)
        nflat = (fw < tol[:, None]).sum(axis=1)
        nneg = (fw < -tol[:, None]).sum(axis=1)
        ok = (nflat == 1) & (nneg == 0)
        ii, jj, fw = ii[ok], jj[ok], fw[ok]
        ndim = np.argmin(fw, axis=1)
        right = origins[jj, ndim] >= origins[ii, ndim]
        out.append(np.stack([ndim[right], ii[right], jj[right]], axis=1))
    faces = np.concatenate(out, axis=0) if out else np.zeros((0, 3), np.int64)
    order = np.lexsort((faces[:, 2], faces[:, 1]))
    return faces[order]


def hcube_faces(hc_origin, hc_widths, origins, widths):
    """``hcube_faces``, ``src/ImmersedBoundary.jl:150-184``."""
    out = []
    for dim in range(len(hc_origin)):
        lo = np.flatnonzero(np.abs(origins[:, dim] - hc_origin[dim]) < widths[:, dim] * F32(0.01))
        out.append(np.stack([np.full(lo.size, dim), np.full(lo.size, -1), lo], axis=1))
        hi = np.flatnonzero(
            np.abs(origins[:, dim] + widths[:, dim] - hc_origin[dim] - hc_widths[dim]) < widths[:, dim] * F32(0.01))
        out.append(np.stack([np.full(hi.size, dim), hi, np.full(hi.size, -1)], axis=1))
    return np.concatenate(out, axis=0).astype(np.int64)


# ------------------------------------------------------------------ ghosts
def ghosts_and_projections_surface(dfield, centers, widths, ghost_layer_ratio=F32(1.5)):
    """Surface version, ``src/ImmersedBoundary.jl:194-230``."""
    glr = F32(ghost_layer_ratio)
    diams = np.sqrt(_sumsq_rows(widths))
    dists = dfield.distances(centers)
    ghosts = np.flatnonzero(dists <= diams * glr * F32(2))
    projs = dfield.projection(centers[ghosts], diams[ghosts] * glr * F32(2)).astype(centers.dtype)
    d = np.sqrt(_sumsq_rows(projs - centers[ghosts]))
    mask = d <= diams[ghosts] * glr
    return ghosts[mask], projs[mask]


def ghosts_and_projections_hcube(faces, hc_origin, hc_widths, centers, widths, ghost_layer_ratio=F32(1.5)):
    """Hypercube version, ``src/ImmersedBoundary.jl:258-305``; faces: [(dim0, front)]."""
    glr = F32(ghost_layer_ratio)
    diams = np.sqrt(_sumsq_rows(widths))
    n = centers.shape[0]
    mask = np.zeros(n, dtype=bool)
    projs = np.zeros_like(centers)
    dists = np.full(n, np.inf, dtype=centers.dtype)
    for dim, front in faces:
        plane = F32(hc_origin[dim] + hc_widths[dim]) if front else F32(hc_origin[dim])
        ps = centers.copy()
        ps[:, dim] = plane
        ds = np.sqrt(_sumsq_rows(ps - centers))
        closer = ds < dists
        dists = np.where(closer, ds, dists)
        projs[closer] = ps[closer]
        mask |= ds < diams * glr
    ghosts = np.flatnonzero(mask)
    return ghosts, projs[ghosts]


class Boundary:
    """``Boundary`` struct + constructor, ``src/ImmersedBoundary.jl:406-448``."""

    def __init__(self, centers, widths, tree, ghost_indices, projs, ghost_ratio=F32(1.5)):
        ghosts = centers[ghost_indices]
        normals = ghosts - projs
        self.ghost_distances = np.sqrt(_sumsq_rows(normals))
        self.normals = normals / (self.ghost_distances + EPS32)[:, None]
        self.image_distances = np.sqrt(_sumsq_rows(widths[ghost_indices])) * F32(ghost_ratio) + EPS32
        self.images = projs + self.normals * self.image_distances[:, None]
        self.ghost_indices = ghost_indices.astype(np.int64)
        self.projections = projs
        idx, w, mask = interpolator_tables(centers, self.images, tree, linear=True)
        self.donor_idx, self.donor_w, self.donor_mask = idx, w, mask  # raw (global) tables
        lens = mask.sum(axis=1)
        ptr = np.concatenate([[0], np.cumsum(lens)])
        interp = Accumulator.from_csr(ptr, idx[mask], w[mask], first_index=True)
        self.image_domain = interp.domain()
        hmap = np.full(centers.shape[0], -1, dtype=np.int64)
        hmap[self.image_domain] = np.arange(self.image_domain.size)
        interp.re_index(hmap)
        self.image_interpolator = interp


def boundary_partitions(centers, widths, tree, ghost_indices, projs, max_partition_size=100_000, ghost_ratio=F32(1.5)):
    """``boundary_partitions``, ``src/ImmersedBoundary.jl:456-476`` -> {ipart: Boundary} (ipart from 1)."""
    bd = {}
    for ipart, s in enumerate(range(0, len(ghost_indices), max_partition_size), start=1):
        sl = slice(s, s + max_partition_size)
        bd[ipart] = Boundary(centers, widths, tree, ghost_indices[sl], projs[sl], ghost_ratio)
    return bd


class Surface:
    """``Surface``, ``src/ImmersedBoundary.jl:335-376``."""

    def __init__(self, points, offsets, normals, areas, interpolator, offset_interpolator, stl):
        self.points, self.offsets, self.normals, self.areas = points, offsets, normals, areas
        self.interpolator, self.offset_interpolator, self.stl = interpolator, offset_interpolator, stl

    def __call__(self, u):
        return self.interpolator(u)

    def at_offset(self, u):
        return self.offset_interpolator(u)


def surface_integral(surf, u):
    """``surface_integral``, ``src/ImmersedBoundary.jl:351-361``."""
    u = np.asarray(u)
    if u.ndim == 1:
        return (surf.areas * u).sum()
    return (surf.areas[:, None] * u).sum(axis=0)


class Partition:
    """``Partition``, ``src/ImmersedBoundary.jl:383-392``."""

    def __init__(self, pid, centers, spacing, face_accumulators, face_owners_neighbors, domain, image, image_in_domain,
                 face_lists=None):
        self.id = pid
        self.centers = centers
        self.spacing = spacing
        self.face_accumulators = face_accumulators
        self.face_owners_neighbors = face_owners_neighbors
        self.domain = domain
        self.image = image
        self.image_in_domain = image_in_domain
        self.face_lists = face_lists  # {(dim, side): (ptr, idx)} CSR copies for table comparison

    @property
    def ndims(self):
        return self.centers.shape[1]


def _group_lists(keys, vals, n):
    """CSR of `vals` grouped by `keys` (stable: append order preserved)."""
    order = np.argsort(keys, kind="stable")
    cnt = np.bincount(keys, minlength=n)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    return ptr, vals[order]


def build_partition(pid, image, faces, c2f_ptr, c2f_idx, centers, widths, skirt_depth=2):
    """One iteration of the partition loop, ``src/ImmersedBoundary.jl:605-703``."""
    ncells, nd = centers.shape
    fo, fn = faces[:, 1], faces[:, 2]
    in_dom = np.zeros(ncells, dtype=bool)
    in_dom[image] = True
    for _ in range(skirt_depth):
        cells = np.flatnonzero(in_dom)
        fidx = c2f_idx[_ranges(c2f_ptr[cells], c2f_ptr[cells + 1])]
        o, nb = fo[fidx], fn[fidx]
        in_dom[o[o >= 0]] = True
        in_dom[nb[nb >= 0]] = True
    domain = np.flatnonzero(in_dom)  # sorted
    idx2domain = np.full(ncells + 1, -1, dtype=np.int64)  # slot ncells catches the "-1" (no cell) lookups
    idx2domain[domain] = np.arange(domain.size)
    # face_indices = reduce(union, cells2faces[domain]) |> unique  (first-appearance order)
    allf = c2f_idx[_ranges(c2f_ptr[domain], c2f_ptr[domain + 1])]
    _, first = np.unique(allf, return_index=True)
    face_indices = allf[np.sort(first)]
    accs, own_nei, lists = {}, {}, {}
    nloc = domain.size
    for dim in range(nd):
        fsel = face_indices[faces[face_indices, 0] == dim]
        o = idx2domain[fo[fsel]]
        nb = idx2domain[fn[fsel]]
        add_right = o >= 0
        o = np.where(add_right, o, nb)
        add_left = nb >= 0
        nb = np.where(add_left, nb, o)
        k = np.arange(fsel.size, dtype=np.int64)
        lptr, lidx = _group_lists(nb[add_left], k[add_left], nloc)
        rptr, ridx = _group_lists(o[add_right], k[add_right], nloc)
        own_nei[dim] = (o, nb)
        for side, (ptr, idx) in ((False, (lptr, lidx)), (True, (rptr, ridx))):
            lens = np.diff(ptr)
            w = (F32(1.0) / np.repeat(lens, lens).astype(F32)).astype(F32)  # _averaging_weights :501-506
            accs[(dim, side)] = Accumulator.from_csr(ptr, idx, w, first_index=True)
            lists[(dim, side)] = (ptr, idx)
    return Partition(pid, centers[domain], widths[domain], accs, own_nei, domain, image, idx2domain[image], lists)


def _ranges(starts, ends):
    """Concatenate arange(s, e) for all (s, e) pairs."""
    lens = ends - starts
    tot = int(lens.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    off = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    return np.arange(tot, dtype=np.int64) + off


class Domain:
    """``Domain`` struct + constructor, ``src/ImmersedBoundary.jl:483-786``."""

    def __init__(self, msh, max_partition_size=100_000, partition_skirt_depth=2, ghost_layer_ratio=F32(1.5),
                 hypercube_families=(), build_surfaces=True):
        self.mesh = msh
        nd = msh.nd
        self.ncells = len(msh)
        centers, widths = get_cells(msh)
        origins = centers - widths / F32(2)
        self.centers, self.widths = centers, widths
        self.faces = np.concatenate(
            [octree2faces(origins, widths), hcube_faces(msh.origin, msh.widths, origins, widths)], axis=0)
        nf = self.faces.shape[0]
        # cells2faces, :575-585: for every face in order push to owner list, then to neighbour list
        cell = np.stack([self.faces[:, 1], self.faces[:, 2]], axis=1).ravel()
        fidx = np.repeat(np.arange(nf, dtype=np.int64), 2)
        ok = cell >= 0
        self.c2f_ptr, self.c2f_idx = _group_lists(cell[ok], fidx[ok], self.ncells)
        self.partitions = {}
        for ipart, s in enumerate(range(0, self.ncells, max_partition_size), start=1):
            image = np.arange(s, min(self.ncells, s + max_partition_size), dtype=np.int64)
            self.partitions[ipart] = build_partition(ipart, image, self.faces, self.c2f_ptr, self.c2f_idx, centers,
                                                     widths, partition_skirt_depth)
        self.boundaries, self.surfaces = {}, {}
        tree = KDTree(centers)
        self.tree = tree
        diams = np.sqrt(_sumsq_rows(widths))
        for bname, hfaces in hypercube_families:
            g, p = ghosts_and_projections_hcube(hfaces, msh.origin, msh.widths, centers, widths, ghost_layer_ratio)
            self.boundaries[bname] = boundary_partitions(centers, widths, tree, g, p, max_partition_size, ghost_layer_ratio)
        for bname, dfield in msh.distance_fields.items():
            g, p = ghosts_and_projections_surface(dfield, centers, widths, ghost_layer_ratio)
            self.boundaries[bname] = boundary_partitions(centers, widths, tree, g, p, max_partition_size, ghost_layer_ratio)
            if not build_surfaces or dfield.stl is None:
                continue
            stl = dfield.stl  # :743-763
            fcenters, fnormals = centers_and_normals(stl)
            idx, _ = tree.nn(fcenters)
            h = diams[idx] * F32(1.01)
            A = (np.sqrt(_sumsq_rows(fnormals)) + EPS32)
            fn_ = fnormals / A[:, None]
            bias = (fn_ * h[:, None])
            self.surfaces[bname] = Surface(
                fcenters, h, fn_, A,
                Interpolator(centers, fcenters, tree, bias=bias, first_index=True),
                Interpolator(centers, fcenters + bias * F32(ghost_layer_ratio), tree, first_index=True), stl)
        self.reconstruction_kwargs = dict(max_partition_size=max_partition_size,
                                          partition_skirt_depth=partition_skirt_depth,
                                          ghost_layer_ratio=ghost_layer_ratio,
                                          hypercube_families=list(hypercube_families))

    @property
    def ndims(self):
        return self.mesh.nd

    def __len__(self):
        return len(self.mesh)

    def __call__(self, f, *args, n_threads=1, **kwargs):
        """``(dom::Domain)(f, args...)``, ``src/ImmersedBoundary.jl:820-864``: gather -> f -> scatter image rows.

        Results come back in ascending partition id (the reference's Dict-key order is arbitrary).
        """
        def run(i):
            part = self.partitions[i]
            dargs = [a[part.domain].copy() for a in args]
            r = f(part, *dargs, **kwargs)
            for a, da in zip(args, dargs):
                a[part.image] = da[part.image_in_domain]
            return r

        keys = sorted(self.partitions)
        if n_threads <= 1 or len(keys) == 1:
            return [run(i) for i in keys]
        with ThreadPoolExecutor(n_threads) as ex:
            return list(ex.map(run, keys))


# ------------------------------------------------------------------ grid operators
def _col(a, v):
    """Broadcast a per-row vector against a (rows,) or (rows, nv) array."""
    return a if v.ndim == 1 else a[:, None]


def at_owners(part, u, dim):
    """``at_owners``, ``src/ImmersedBoundary.jl:879-881``."""
    return u[part.face_owners_neighbors[dim][0]]


def at_neighbors(part, u, dim):
    """``at_neighbors``, ``src/ImmersedBoundary.jl:889-891``."""
    return u[part.face_owners_neighbors[dim][1]]


def at_faces(part, u, dim):
    """``at_faces``, ``src/ImmersedBoundary.jl:899-910``."""
    spo = at_owners(part, part.spacing, dim)[:, dim]
    spn = at_neighbors(part, part.spacing, dim)[:, dim]
    uo, un = at_owners(part, u, dim), at_neighbors(part, u, dim)
    return (uo * _col(spn, uo) + un * _col(spo, uo)) / _col(spn + spo, uo)


def green_gauss(part, uf, dim):
    """``green_gauss``, ``src/ImmersedBoundary.jl:918-926``."""
    accl, accr = part.face_accumulators[(dim, False)], part.face_accumulators[(dim, True)]
    return (accr(uf) - accl(uf)) / _col(part.spacing[:, dim], uf)


def unsigned_green_gauss(part, uf, dim):
    """``unsigned_green_gauss``, ``src/ImmersedBoundary.jl:934-942``."""
    accl, accr = part.face_accumulators[(dim, False)], part.face_accumulators[(dim, True)]
    return (accr(uf) + accl(uf)) / _col(part.spacing[:, dim], uf)


def divergent(part, uf):
    """``divergent``, ``src/ImmersedBoundary.jl:950-956``."""
    s = green_gauss(part, uf[0], 0)
    for dim in range(1, part.ndims):
        s = s + green_gauss(part, uf[dim], dim)
    return s


def cell_gradient(part, u, dim=None):
    """``cell_gradient``, ``src/ImmersedBoundary.jl:965-987``."""
    if dim is None:
        return tuple(cell_gradient(part, u, d) for d in range(part.ndims))
    return green_gauss(part, at_faces(part, u, dim), dim)


def face_distance(part, dim):
    """``face_distance``, ``src/ImmersedBoundary.jl:995-1002``."""
    return (at_owners(part, part.spacing, dim)[:, dim] + at_neighbors(part, part.spacing, dim)[:, dim]) / F32(2)


def owner_distance(part, dim):
    """``owner_distance``, ``src/ImmersedBoundary.jl:1010-1016``."""
    return at_owners(part, part.spacing, dim)[:, dim] / F32(2)


def neighbor_distance(part, dim):
    """``neighbor_distance``, ``src/ImmersedBoundary.jl:1024-1030``."""
    return at_neighbors(part, part.spacing, dim)[:, dim] / F32(2)


def face_gradient(part, u, dim, grad_u=None):
    """``face_gradient``, ``src/ImmersedBoundary.jl:1039-1069``."""
    if grad_u is None:
        d = at_neighbors(part, u, dim) - at_owners(part, u, dim)
        return d / _col(face_distance(part, dim), d)
    return tuple(face_gradient(part, u, dim) if i == dim else at_faces(part, grad_u[i], dim)
                 for i in range(part.ndims))


def JST_sensor(part, p, dim=None):
    """``CFD.JST_sensor(part, p, dim)``, ``src/ImmersedBoundary.jl:1077-1097`` (dim None == reference dim 0)."""
    if dim is None:
        nu = np.full_like(p, F32(1e-7))
        for d in range(part.ndims):
            nu = np.maximum(nu, JST_sensor(part, p, d))
        return nu
    fd = at_neighbors(part, p, dim) - at_owners(part, p, dim)
    return (F32(1e-7) + np.abs(green_gauss(part, fd, dim))) / (F32(1e-7) + unsigned_green_gauss(part, np.abs(fd), dim))


def minmod(a, b):
    """``minmod``, ``src/ImmersedBoundary.jl:1099``."""
    return np.minimum(np.abs(a), np.abs(b)) * (np.sign(a) + np.sign(b)) / 2


def MUSCL(part, u, du, dim, D=None, high_order=False):
    """``MUSCL``, ``src/ImmersedBoundary.jl:1113-1157``."""
    down, dnei = owner_distance(part, dim), neighbor_distance(part, dim)
    uo, un = at_owners(part, u, dim), at_neighbors(part, u, dim)
    c = lambda a: _col(a, uo)
    gf = (un - uo) / c(down + dnei)
    duo, dun = at_owners(part, du, dim), at_neighbors(part, du, dim)
    gu = (2 * duo - gf) * c(down)
    Du = (2 * dun - gf) * c(dnei)
    s = minmod(Du, gu)
    uL, uR = uo + s, un - s
    if D is not None:
        Df = np.maximum(np.maximum(at_owners(part, D, dim), at_neighbors(part, D, dim)), F32(1e-7))
        uf = (uo * c(dnei) + un * c(down)) / c(down + dnei)
        if high_order:
            uf = uf + (duo * c(down) - dun * c(dnei)) / 8
        uL = uL * c(Df) + (F32(1.0) - c(Df)) * uf
        uR = uR * c(Df) + (F32(1.0) - c(Df)) * uf
    return uL, uR


# ------------------------------------------------------------------ IB ghost update
def impose_bc(f, dom, bname, *args, **kwargs):
    """``impose_bc!``, ``src/ImmersedBoundary.jl:1197-1247``.

    ``f(bdry, *image_values)`` may return a scalar, an array or a tuple; ``zip``
    truncation makes trailing args auxiliary.  All boundary partitions read before any
    of them writes (the reference races across partitions; with one partition -- every
    shipped case -- the two are identical).
    """
    parts = dom.boundaries[bname]
    pending = []
    for ipart in sorted(parts):
        bdry = parts[ipart]
        eta = bdry.ghost_distances / bdry.image_distances
        iargs = [bdry.image_interpolator(a[bdry.image_domain]) for a in args]
        r = f(bdry, *iargs, **kwargs)
        if not isinstance(r, tuple):
            r = (r,)
        for a, ba, ia in zip(args, r, iargs):
            e = _col(eta, ia)
            pending.append((a, bdry.ghost_indices, e * ia + (F32(1.0) - e) * ba))
    for a, g, val in pending:  # Jacobi: every image read above precedes every ghost write
        a[g] = val


# ------------------------------------------------------------------ multigrid / reductions
def multigrid(dom, max_levels=0, factor=2):
    """``multigrid``, ``src/ImmersedBoundary.jl:1355-1407``.

    Returns ``(coarse_doms, prolongators, coarseners)`` -- the order the reference
    *code* returns (``:1406``), not the order its docstring states.
    """
    msh = dom.mesh
    mdepth = int(np.floor(np.log2(msh.block_size)))
    max_levels = mdepth if max_levels == 0 else max_levels
    coarse_doms, coarseners, prolongators = [], [], []
    Xold, tree_old = dom.centers, KDTree(dom.centers)
    bsize = msh.block_size
    for _ in range(max_levels):
        bsize //= factor
        cm = Mesh.from_blocks(msh.origin, msh.widths, bsize, msh.block_origins, msh.block_widths, msh.distance_fields)
        cdom = Domain(cm, **dom.reconstruction_kwargs)
        X = cdom.centers
        tree = KDTree(X)
        coarseners.append(Interpolator(Xold, X, tree_old, first_index=True, linear=False))
        prolongators.append(Interpolator(X, Xold, tree, first_index=True, linear=False))
        coarse_doms.append(cdom)
        Xold, tree_old = X, tree
    return coarse_doms, prolongators, coarseners


def volume_integral(dom, A):
    """``volume_integral``, ``src/ImmersedBoundary.jl:1415-1431``."""
    Ai = np.array(A, dtype=F32, copy=True)

    def f(part, Ai):
        for dim in range(part.ndims):
            Ai *= _col(part.spacing[:, dim], Ai)

    dom(f, Ai)
    return Ai.sum(axis=0, dtype=np.float32)
