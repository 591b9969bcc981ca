"""Diagnostic (GPU): localise differences between the fused Euler paths and the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import immersedboundary_jl_b200 as ib
import oracle
from conftest import Case
F32 = np.float32
name = sys.argv[1] if len(sys.argv) > 1 else "sphere3d"
c = Case(name, ib, oracle, 40_000, upload=True)
fl, ofl = ib.Fluid(), oracle.cfd.Fluid()
N, nd = len(c.dom), c.dom.ndims
Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c.odom.centers))
Ro, co = np.zeros_like(Q0), np.zeros(N, F32)
c.odom(oracle.euler.euler_residual(ofl), Q0.copy(), Ro, co)
scale = np.abs(Ro).max(axis=0)
bf = c.dom.block_faces()
cpb = 8 ** nd
Q = ib.DeviceArray.from_host(Q0)
for label, generic in (("tile", False), ("generic", True)):
    if generic:
        os.environ["IBX_GENERIC"] = "1"
    R, cf = ib.DeviceArray(N, nd + 2, False), ib.DeviceArray(N, 1, True)
    ib.residual_euler(c.dom, fl, Q, R, cf)
    os.environ.pop("IBX_GENERIC", None)
    Rg, cg = R.to_host(), cf.to_host()
    err = np.abs(Rg - Ro) / scale
    e = err.max(axis=1)
    bad = np.flatnonzero(e > 1e-5)
    ndiff = int((Rg != Ro).sum())
    ulp = np.abs(Rg.view(np.int32).astype(np.int64) - Ro.astype(F32).view(np.int32).astype(np.int64))
    print(f"== {label}: max scaled err {e.max():.3e}; cells with err > 1e-5: {len(bad)} of {N}; cfl max rel err {np.abs(cg / co - 1).max():.3e}")
    print(f"   values that differ at all: {ndiff} of {Rg.size} ({ndiff / Rg.size:.2e}); largest difference {int(ulp.max())} ulp")
    if len(bad):
        blk = bad // cpb
        loc = bad % cpb
        ii = np.stack([loc % 8, (loc // 8) % 8, loc // 64 if nd == 3 else 0 * loc], axis=1)
        kinds = bf[blk][:, :, 0]
        print("   bad cells by block-face kinds present (box, same, coarser, finer):",
              [(int((kinds == k).any(axis=1).sum())) for k in range(4)])
        on_face = ((ii[:, :nd] == 0) | (ii[:, :nd] == 7)).any(axis=1)
        near_face = ((ii[:, :nd] <= 1) | (ii[:, :nd] >= 6)).any(axis=1)
        print("   on block face:", int(on_face.sum()), " within 2 of a face:", int(near_face.sum()), " interior:", int((~near_face).sum()))
        worst = bad[np.argsort(-e[bad])[:8]]
        for w in worst:
            b, l = w // cpb, w % cpb
            print("   cell", w, "block", b, "local", (l % 8, (l // 8) % 8, l // 64), "err", err[w], "kinds", bf[b][:, 0].tolist())
            print("      R gpu", Rg[w], "\n      R ref", Ro[w])
