"""C3 (RAE2822, M = 0.73, alpha = 2.31 deg, Euler) marched to its steady state on the device through `ib.march_euler`;
prints / writes the lift and drag history next to the oracle's converged values of tests/golden/rae2822_converged.npz.

    python tools/c3_gpu_march.py --out gpurun_out/c3_polar.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F32 = np.float32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=0, help="0: as many as the fixture's march")
    ap.add_argument("--every", type=int, default=5000)
    args = ap.parse_args()
    import immersedboundary_jl_b200 as ib
    g = np.load(os.path.join(ROOT, "tests", "golden", "rae2822_converged.npz"))
    steps = args.steps or int(g["steps"])
    cfl, stages, mach, alpha = F32(g["cfl"]), int(g["stages"]), float(g["mach"]), float(g["alpha"])
    RAE = os.path.join(ROOT, "tests", "golden", "rae2822.dat")
    stl = ib.merge_points(ib.Stereolitography(RAE))
    feat = ib.DistanceField(ib.feature_regions(stl, radius=0.05))
    msh = ib.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", stl, F32(1e-2)), refinement_regions=[(feat, F32(5e-3))])
    fams = [("farfield", [(0, False), (0, True), (1, False), (1, True)])]
    dom = ib.Domain(msh, max_partition_size=10_000, hypercube_families=fams, upload=True)
    fl = ib.Fluid()
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    al = np.radians(alpha)
    Pinf = np.array([101325.0, 288.15, mach * a_inf * np.cos(al), mach * a_inf * np.sin(al)], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    N = len(dom)
    P0 = np.tile(Pinf, (N, 1))
    P0[ib.synthetic.inside_polygon(np.loadtxt(RAE), dom.cells()[0]), 2:] = 0
    Q = ib.DeviceArray.from_host(np.asfortranarray(ib.synthetic.primitive2state_host(P0)))
    ghost = np.zeros(N, bool)
    for chunks in dom.boundaries.values():
        for b in chunks.values():
            ghost[b.ghost_indices] = True
    live = ib.DeviceArray.from_host((~ghost).astype(F32))
    s = dom.surfaces["wall"]

    def lift_drag(Q):
        p = ib.state2primitive(fl, Q).col(0)
        Cp_s = s(ib.pressure_coefficient(fl, p, Pinf[0], mach)).to_host().ravel()
        F = ib.surface_integral(s, np.asfortranarray(Cp_s[:, None] * s.normals))
        return float(-F[0] * np.sin(al) + F[1] * np.cos(al)), float(F[0] * np.cos(al) + F[1] * np.sin(al))

    hist = []
    t0 = time.time()

    def monitor(it, Q, R, cf):
        cl, cd = lift_drag(Q)
        rn = float((R.col(0) / cf * live).norm())
        hist.append({"step": it, "cl": cl, "cd": cd, "res_rho": rn, "t": round(time.time() - t0, 2)})
        print(hist[-1], flush=True)

    ib.march_euler(dom, fl, bcs, Q, steps, CFL=cfl, stages=stages, live=live, monitor=monitor, every=args.every)
    cl, cd = lift_drag(Q)
    Qg, Qo = Q.to_host(), g["Q"]
    err = (np.abs(Qg - Qo) / np.abs(Qo).max(axis=0)).max(axis=1)
    out = {"case": "rae2822 M=0.73 alpha=2.31 Euler", "cells": N, "steps": steps, "stages": stages, "cfl": float(cfl),
           "seconds": round(time.time() - t0, 2), "gpu": {"cl": cl, "cd": cd}, "oracle": {"cl": float(g["cl"]), "cd": float(g["cd"])},
           "abs_diff": {"cl": abs(cl - float(g["cl"])), "cd": abs(cd - float(g["cd"]))},
           "state_err_of_scale": {"mean": float(err.mean()), "max": float(err.max())} if steps == int(g["steps"]) else None,
           "history": hist}
    print(json.dumps({k: v for k, v in out.items() if k != "history"}))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
