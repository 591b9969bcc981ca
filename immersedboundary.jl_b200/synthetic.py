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
