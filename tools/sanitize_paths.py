"""Driver for compute-sanitizer runs (racecheck / initcheck / memcheck) of the product paths the tests exercise:
the 2-D host-buffer march of C3 (ibx_euler_step_host, the path of the round-1 flake), the two-slot pipelined form on a
3-D mesh, and the multigrid-level kernels (block sizes 4 and 2).  Product only: no oracle, small meshes.

    compute-sanitizer --tool racecheck python tools/sanitize_paths.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import immersedboundary_jl_b200 as ib  # noqa: E402

F32 = np.float32
fl = ib.Fluid()
a_inf = np.sqrt(1.4 * 283.0 * 288.15)
wall = np.array([101325.0, 288.15, 0.0], F32)

# ---- 2-D: RAE2822, host-buffer march
RAE = os.path.join(ROOT, "tests", "golden", "rae2822.dat")
stl = ib.merge_points(ib.Stereolitography(RAE))
feat = ib.DistanceField(ib.feature_regions(stl, radius=0.05))
msh = ib.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", stl, F32(1e-2)), refinement_regions=[(feat, F32(5e-3))])
dom = ib.Domain(msh, max_partition_size=10_000, hypercube_families=[("farfield", [(0, False), (0, True), (1, False), (1, True)])])
d = ib.streamwise_direction(2.31)
Pinf = np.array([101325.0, 288.15, 0.73 * a_inf * d[0], 0.73 * a_inf * d[1]], F32)
bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
N = len(dom)
Q = ib.pinned_empty((N, 4))
Q[...] = ib.synthetic.primitive2state_host(np.tile(Pinf, (N, 1)))
R, cf = ib.pinned_empty((N, 4)), ib.pinned_empty((N,))
for _ in range(3):
    ib.euler_step_host(dom, fl, bcs, Q, R, cf)
    Q += (F32(0.4) / cf)[:, None] * R
print("2-D host march ok", float(np.abs(R).max()))
# multigrid levels of the same domain (block sizes 4 and 2)
cd, pro, coa = ib.multigrid(dom)
for lvl in cd[:2]:
    n = len(lvl)
    Ql = ib.DeviceArray.from_host(ib.synthetic.primitive2state_host(np.tile(Pinf, (n, 1))))
    Rl, cl = ib.DeviceArray(n, 4, False), ib.DeviceArray(n, 1, True)
    ib.ghost_update_euler(lvl, fl, Ql, bcs)
    ib.residual_euler(lvl, fl, Ql, Rl, cl)
    for flux in ("hll", "sensor"):
        ib.residual_euler(lvl, fl, Ql, Rl, cl, flux=flux)
print("2-D multigrid levels ok")

# ---- 3-D: sphere, two slots in flight
fams = [("farfield", [(dd, s) for dd in range(3) for s in (False, True)])]
m3 = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12)), refinement_regions=[(ib.Ball([0, 0, 0], 0.9), F32(0.12))])
d3 = ib.Domain(m3, hypercube_families=fams, build_partitions=False)
N3 = len(d3)
P3 = np.array([101325.0, 288.15, 0.5 * a_inf, 0.0, 0.0], F32)
b3 = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, P3))]
Qp = [ib.pinned_empty((N3, 5)) for _ in range(2)]
for q in Qp:
    q[...] = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(d3.cells()[0]))
Rs, cs = [ib.pinned_empty((N3, 5)) for _ in range(2)], [ib.pinned_empty((N3,)) for _ in range(2)]
for rep in range(2):
    for k in range(2):
        ib.euler_step_host_begin(d3, fl, b3, Qp[k], Rs[k], cs[k], k)
    for k in range(2):
        ib.euler_step_host_end(k)
assert np.array_equal(Rs[0], Rs[1])
for flux in ("hll", "sensor"):
    Qd = ib.DeviceArray.from_host(Qp[0])
    Rd, cd_ = ib.DeviceArray(N3, 5, False), ib.DeviceArray(N3, 1, True)
    ib.residual_euler(d3, fl, Qd, Rd, cd_, flux=flux)
print("3-D two-slot + both fluxes ok")
ib.synchronize()
