"""GPU check + timing: pencil-marching flux kernel (march.cu) against the tile kernels (tile.cu) on the C4 mesh.

    python tools/march_check.py [level] [radius]     # level 10 / radius 0.75 = the 50.2 M-cell bench mesh
Both paths claim the same bits; this prints the number of cells whose residual / CFL differ and the time per call."""
import ctypes as C
import math
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import immersedboundary_jl_b200 as ib
from bench import build_mesh
F32 = np.float32
level = int(sys.argv[1]) if len(sys.argv) > 1 else 10
radius = float(sys.argv[2]) if len(sys.argv) > 2 else 0.75
flux = sys.argv[3] if len(sys.argv) > 3 else "hll"
ctx = ib.context(0)
msh = build_mesh(ib, radius, 32.0 / 2 ** level / 8 * 1.01)
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
print(f"cells {N}  blocks {msh.nblocks}  flux {flux}", flush=True)
ms = C.c_float()
ref = None
for label, env in (("tile kernels", {"IBX_NO_MARCH": "1"}), ("march, generic general faces", {"IBX_GEN_OLD": "1"}),
                   ("march SEG=1", {"IBX_MARCH_SEG": "1"}), ("march scalar FADD/FMUL", {"IBX_MARCH_SCALAR": "1"}),
                   ("march, middle face twice", {"IBX_MARCH_NOSHARE": "1"}), ("march (default)", {}),
                   ("march, HLL on (L, R) pairs", {"IBX_MARCH_HLR": "1"}), ("march (default) again", {})):
    for k in ("IBX_NO_MARCH", "IBX_MARCH_SEG", "IBX_GEN_OLD", "IBX_MARCH_SCALAR", "IBX_MARCH_NOSHARE", "IBX_MARCH_HLR"):
        os.environ.pop(k, None)
    os.environ.update(env)
    R.fill(0.0); cfl.fill(0.0)
    for _ in range(2):
        ib.residual_euler(dom, fluid, Q, R, cfl, flux=flux)
    reps = 10
    ib._lib.call("ibx_timer_start", ctx)
    for _ in range(reps):
        ib.residual_euler(dom, fluid, Q, R, cfl, flux=flux)
    ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
    Rh, ch = R.to_host(), cfl.to_host()
    msg = f"{label:30s} {ms.value / reps:8.3f} ms/call  {N / (ms.value / reps) / 1e6:8.2f} G cell-updates/s"
    if ref is None:
        ref = (Rh, ch)
    else:
        dr = (Rh != ref[0]).any(axis=1)
        dc = ch != ref[1]
        msg += f"   cells with different R: {int(dr.sum())}, different cfl: {int(dc.sum())}"
        if dr.any():
            scale = np.abs(ref[0]).max(axis=0)
            msg += f", max scaled diff {(np.abs(Rh - ref[0]) / scale).max():.3e}, nan {int(np.isnan(Rh).sum())}"
            bad = np.flatnonzero(dr)[:5]
            for w in bad:
                b, l = w // 512, w % 512
                msg += f"\n   cell {w} block {b} local {(l % 8, (l // 8) % 8, l // 64)}: {Rh[w]} vs {ref[0][w]}"
    print(msg, flush=True)
