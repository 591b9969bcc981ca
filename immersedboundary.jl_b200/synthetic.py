"""Synthetic inputs for tests and benchmarks: icosphere STL, sphere octree meshes,
and the smooth Euler state.  Pure NumPy; shared by the product-side bench and the parity tests."""
import numpy as np

F32 = np.float32


def icosphere(subdivisions=2, radius=0.5, center=(0.0, 0.0, 0.0)):
    """Unit icosahedron subdivided `subdivisions` times -> (points (n, 3) float32, triangles (m, 3) int64)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    for _ in range(subdivisions):
        cache = {}
        verts = list(v)

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (verts[a] + verts[b]) / 2
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.asarray(verts), np.asarray(nf, dtype=np.int64)
    return (v * radius + np.asarray(center)).astype(F32), f


def sphere_regions(h_wall, radii_and_sizes):
    return [(r, F32(h)) for r, h in radii_and_sizes]


def euler_state(centers, fluid_R=283.0, gamma=1.4, mach=0.5, seed=12345, p_inf=101325.0, T_inf=288.15):
    """Smooth primitive field of SURVEY.md section 8(d) + 1e-4 relative noise -> P (N, 2 + nd) float32."""
    x = centers.astype(np.float64)
    nd = x.shape[1]
    k = 2 * np.pi / 8.0
    a_inf = np.sqrt(gamma * fluid_R * T_inf)
    u_inf = mach * a_inf
    z = x[:, 2] if nd > 2 else np.zeros(len(x))
    p = p_inf * (1 + 0.02 * np.sin(k * x[:, 0]) * np.cos(k * x[:, 1]) * np.cos(k * z))
    T = T_inf * (1 + 0.01 * np.cos(k * x[:, 0]))
    cols = [p, T, u_inf * (1 + 0.05 * np.sin(k * x[:, 1])), 0.05 * u_inf * np.sin(k * z if nd > 2 else k * x[:, 0])]
    if nd > 2:
        cols.append(0.05 * u_inf * np.sin(k * x[:, 0]))
    P = np.stack(cols, axis=1)
    rng = np.random.default_rng(seed)
    P *= 1 + 1e-4 * (rng.random(P.shape) * 2 - 1)
    return P.astype(F32)


def primitive2state_host(P, R=283.0, gamma=1.4):
    """float32 NumPy ``primitive2state`` (``src/cfd.jl:106-123``) for building synthetic inputs."""
    P = P.astype(F32)
    R, gamma = F32(R), F32(gamma)
    T = np.maximum(P[:, 1], F32(10.0))
    u = P[:, 2:]
    k = u[:, 0] ** 2
    for d in range(1, u.shape[1]):
        k = k + u[:, d] ** 2
    k = k / F32(2)
    rho = P[:, 0] / (R * T)
    E = rho * (R / (gamma - F32(1.0)) * T + k)
    return np.concatenate([rho[:, None], E[:, None], rho[:, None] * u], axis=1).astype(F32)


def probe_signs(global_ids, nv, n_samples, seed=0):
    """+-1 Hutchinson probes from a counter-based generator keyed by (seed, sample, variable, GLOBAL cell id) -- a
    splitmix64 finaliser of the key -- so that the probes of a cell do not depend on how the mesh is sharded over ranks
    (the reference draws them from the global RNG stream, `src/point_implicit.jl:36`, which cannot be
    split).  Returns float32 (n_samples, nv, len(global_ids))."""
    g = np.asarray(global_ids, dtype=np.uint64)
    out = np.empty((n_samples, nv, g.size), dtype=np.float32)
    with np.errstate(over="ignore"):
        for k in range(n_samples):
            for v in range(nv):
                z = g + np.uint64(0x9E3779B97F4A7C15) * np.uint64(1 + seed + 1000003 * (k * nv + v))
                z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
                z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
                z = z ^ (z >> np.uint64(31))
                out[k, v] = np.where((z >> np.uint64(63)) == 0, np.float32(1.0), np.float32(-1.0))
    return out


def swept_wing(n_chord=64, n_span=64, chord=1.0, span=4.0, sweep_deg=30.0, thickness=0.12):
    """Procedural wing of SURVEY.md 8(d) for configuration C5: a NACA 00xx section (closed trailing edge) extruded along z
    over `span` with `sweep_deg` of sweep, closed by two tip caps -> (points (n, 3) float32, triangles (m, 3) int64).
    The wing is centred on z = 0; the root leading edge sits at x = -span / 2 * tan(sweep)."""
    beta = np.linspace(0.0, np.pi, n_chord + 1)
    x = 0.5 * (1.0 - np.cos(beta))                               # cosine spacing, x[0] = 0 (LE), x[-1] = 1 (TE)
    y = 5.0 * thickness * (0.2969 * np.sqrt(x) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1036 * x ** 4)
    y[0] = y[-1] = 0.0
    # section loop: TE -> upper -> LE -> lower -> (TE): 2 * n_chord distinct points
    sx = np.concatenate([x[::-1], x[1:-1]]) * chord
    sy = np.concatenate([y[::-1], -y[1:-1]]) * chord
    m = len(sx)
    zs = np.linspace(-span / 2, span / 2, n_span + 1)
    tan_s = np.tan(np.radians(sweep_deg))
    pts = np.concatenate([np.stack([sx + z * tan_s, sy, np.full(m, z)], axis=1) for z in zs])
    tri = []
    for k in range(n_span):
        a, b = k * m, (k + 1) * m
        for i in range(m):
            j = (i + 1) % m
            tri += [[a + i, a + j, b + j], [a + i, b + j, b + i]]
    # tip caps: strips between the upper point U_k and the lower point L_k of the same chord station
    up = lambda k: n_chord - k                                   # index of U_k in the loop (U_n = TE at 0, U_0 = LE at n_chord)
    lo = lambda k: n_chord + k if 0 < k < n_chord else up(k)     # L_k (L_0 = LE, L_n = TE coincide with U_0, U_n)
    for base, flip in ((0, False), (n_span * m, True)):
        for k in range(n_chord):
            quad = [base + up(k), base + up(k + 1), base + lo(k + 1), base + lo(k)]
            ts = [[quad[0], quad[1], quad[2]], [quad[0], quad[2], quad[3]]]
            for t in ts:
                if len(set(t)) == 3:
                    tri.append(t[::-1] if flip else t)
    return pts.astype(F32), np.asarray(tri, dtype=np.int64)


def inside_polygon(poly, pts):
    """Even-odd ray casting: which of ``pts`` (n, 2) lie inside the closed loop ``poly`` (m, 2)?  Used to start the cells
    enclosed by a 2-D body at rest (they form a closed cavity the immersed boundary never flushes)."""
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    inside = np.zeros(len(pts), bool)
    n = len(poly)
    for i in range(n):
        (x0, y0), (x1, y1) = poly[i], poly[(i + 1) % n]
        if y0 == y1:
            continue
        cond = (y0 > y) != (y1 > y)
        xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
        inside ^= cond & (x < xi)
    return inside
