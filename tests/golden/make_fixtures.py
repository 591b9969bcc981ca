"""Regenerate tests/golden/oracle_fixtures.npz from the NumPy oracle (run from the repo root).

These are NOT outputs of the Julia reference (it cannot run in this image); they pin the oracle against drift and give
the GPU box reference arrays for configuration C1 (test/advection.jl) and a small 3-D sphere case."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from oracle import mesher as M, domain as D, euler as E, cfd
from immersedboundary_jl_b200 import synthetic as S  # pure NumPy helpers

F32 = np.float32
out = {}
seg = lambda a, b: M.Stereolitography(np.array([a, b], dtype=np.float64))
h = F32(1e-2)
msh = M.Mesh([0.0, 0.0], [1.0, 1.0], ("lower", seg([0., 0.], [1., 0.]), h), ("upper", seg([0., 0.], [0., 1.]), h),
             refinement_regions=[(M.Line([0.0, 0.0], [1.0, 1.0]), F32(2) * h), (M.Line([0.0, 0.0], [0.5, 0.5]), h)])
dom = D.Domain(msh, hypercube_families=[("outlet", [(0, True), (1, True)])])
out["adv_block_origins"], out["adv_block_widths"] = msh.block_origins, msh.block_widths
out["adv_faces"] = dom.faces.astype(np.int32)
for name, bs in dom.boundaries.items():
    out[f"adv_ghosts_{name}"] = bs[1].ghost_indices.astype(np.int32)
rng = np.random.default_rng(7)
u = rng.random(len(dom)).astype(F32)
C = (0.5 + rng.random((len(dom), 2))).astype(F32)
ud = np.zeros(len(dom), F32)
dom(lambda p, u_, ud_, Cl: E.advection_residual(p, u_, ud_, Cl), u.copy(), ud, C.copy())
out["adv_ud"] = ud  # inputs are regenerated from default_rng(7) by the tests

fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
m3 = M.Mesh([-2, -2, -2], [4, 4, 4], ("wall", M.AnalyticSphere([0, 0, 0], 0.5), F32(0.12)))
d3 = D.Domain(m3, hypercube_families=fams)
fl = cfd.Fluid()
Q = S.primitive2state_host(S.euler_state(d3.centers))
R, cf = np.zeros_like(Q), np.zeros(len(Q), F32)
d3(E.euler_residual(fl), Q.copy(), R, cf)
out["sph_block_origins"], out["sph_block_widths"] = m3.block_origins, m3.block_widths
# inputs are regenerated with synthetic.euler_state; outputs: every 16th cell + float64 checksums of all cells
out["sph_R_sub"], out["sph_cfl_sub"] = R[::16].copy(), cf[::16].copy()
out["sph_R_sum"], out["sph_R_abs"] = R.astype(np.float64).sum(axis=0), np.abs(R.astype(np.float64)).sum(axis=0)
out["sph_cfl_sum"] = np.float64(cf.astype(np.float64).sum())
out["sph_ghosts_wall"] = d3.boundaries["wall"][1].ghost_indices.astype(np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.npz"), **out)
print({k: v.shape for k, v in out.items()})

# ---- inputs for tools/gen_reference_fixtures.jl (the Julia pinning kit): the icosphere the 3-D STL tests use
if "--icosphere" in sys.argv:
    pts, tri = S.icosphere(1, 0.5)
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference")
    os.makedirs(ref, exist_ok=True)
    np.ascontiguousarray(pts, dtype=np.float32).tofile(os.path.join(ref, "icosphere1_points.f32"))           # npts x 3 C-order == 3 x npts Julia
    np.ascontiguousarray(tri + 1, dtype=np.int64).tofile(os.path.join(ref, "icosphere1_triangles.i64"))     # 1-based
