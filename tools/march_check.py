"""GPU check + timing on the C4 mesh (or a coarser level of the same recipe): the default residual path against the two
independent implementations behind ibx_set_option("path"), and the fast arithmetic against the exact one.

    python tools/march_check.py [level] [radius] [hll|sensor] [--analytic]   # level 10 / radius 0.75 = the 50.2 M-cell bench mesh
Prints the number of cells whose residual / CFL differ from the default path, the time per call of every variant and
the error of arithmetic = 1 under both normalisations (flux-scaled; residual-scaled)."""
import ctypes as C
import math
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import immersedboundary_jl_b200 as ib
from bench import build_mesh, flux_scaled_error
F32 = np.float32
args = [a for a in sys.argv[1:] if not a.startswith("--")]
level = int(args[0]) if len(args) > 0 else 10
radius = float(args[1]) if len(args) > 1 else 0.75
flux = args[2] if len(args) > 2 else "hll"
ctx = ib.context(0)
msh = build_mesh(ib, radius, 32.0 / 2 ** level / 8 * 1.01, analytic="--analytic" in sys.argv)
fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
dom = ib.Domain(msh, max_partition_size=len(msh), hypercube_families=fams, build_partitions=False, build_surfaces=False, upload=True)
N = len(dom)
fluid = ib.Fluid()
a_inf = math.sqrt(1.4 * 283.0 * 288.15)
Pinf = np.array([101325.0, 288.15, 0.5 * a_inf, 0.0, 0.0], F32)
bcs = [("wall", ib.FlowBC(fluid, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)), ("farfield", ib.FlowBC(fluid, Pinf))]
Q = ib.DeviceArray.from_host(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(dom.cells()[0])))
ib.ghost_update_euler(dom, fluid, Q, bcs)
R, cfl = ib.DeviceArray(N, 5, False), ib.DeviceArray(N, 1, True)
ms = C.c_float()


def run(label, reps=10, **opts):
    with ib.options(**opts):
        for _ in range(2):
            ib.residual_euler(dom, fluid, Q, R, cfl, flux=flux)
        ib._lib.call("ibx_timer_start", ctx)
        for _ in range(reps):
            ib.residual_euler(dom, fluid, Q, R, cfl, flux=flux)
        ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
    print(f"{label:34s} {ms.value / reps:8.3f} ms / call  {N / (ms.value / reps * 1e-3) / 1e9:7.2f} G cells/s", flush=True)
    return R.to_host(), cfl.to_host()


print(f"{N} cells, level {level}, flux {flux}")
R0, c0 = run("default (marching kernels)")
for label, opts in (("path = 1 (tile kernels)", dict(path=1)), ("path = 2 (gather kernels)", dict(path=2))):
    if N > 12_000_000 and opts["path"] == 2:
        continue
    R1, c1 = run(label, reps=3, **opts)
    print(f"    cells differing from default: R {int((R1 != R0).any(axis=1).sum())}, cfl {int((c1 != c0).sum())}")
if flux == "hll":
    R1, c1 = run("arithmetic = 1 (fast)", arithmetic=1)
    Qh = Q.to_host()
    widths = dom.cells()[1]
    e_flux = flux_scaled_error(widths, Qh, R1, R0)
    scale = np.abs(R0).max(axis=0)
    e_res = np.abs(R1 - R0) / np.maximum(np.abs(R0), 1e-3 * scale)
    print(f"    fast vs exact: flux-scaled max {e_flux.max():.3e} (99.9 % {np.quantile(e_flux, 0.999):.3e}); "
          f"residual-scaled max {e_res.max():.3e} (median {np.median(e_res):.3e}); cfl max rel {np.abs(c1 / c0 - 1).max():.2e}")
R1, c1 = run("default again")
print(f"    bit-reproducible: {np.array_equal(R1, R0) and np.array_equal(c1, c0)}")
