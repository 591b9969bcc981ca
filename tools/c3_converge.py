"""CPU experiment (oracle side only): which driver converges the RAE2822 Euler case (BASELINE.json configs[2])?

Marches the restated reference residual (oracle/cpu_ref.c, bit-identical to the NumPy oracle) with local time steps:
    Q <- Q + w * R / cfl         (what `FAS!` does with f(l, Q) = (R * CFL / cfl, 1), src/solver.jl:78-82)
either as a single forward-Euler stage or as an m-stage Runge-Kutta smoother inside f, with or without the JST sensor
blend of MUSCL (`D = nothing`, src/ImmersedBoundary.jl:1141).  Prints the residual norm and Cl / Cd every `--every` steps.

    python tools/c3_converge.py --stages 3 --cfl 1.0 --no-sensor --steps 4000
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F32 = np.float32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--stages", type=int, default=1)
    ap.add_argument("--cfl", type=float, default=0.4)
    ap.add_argument("--no-sensor", action="store_true")
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--mach", type=float, default=0.73)
    ap.add_argument("--alpha", type=float, default=2.31)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--glr", type=float, default=1.5, help="ghost_layer_ratio of Domain (src/ImmersedBoundary.jl:536)")
    ap.add_argument("--ramp", type=int, default=0, help="ramp the free-stream Mach number from 20 % to 100 % over this many steps")
    ap.add_argument("--save", default="", help="write the final state, Cl and Cd to this .npz")
    ap.add_argument("--load", default="", help="continue from the state saved by an earlier run")
    ap.add_argument("--evolve-ghosts", action="store_true", help="let the residual advance the ghost cells too (default: ghost cells only take BC values)")
    ap.add_argument("--copy-ghosts", action="store_true", help="ghost update on a copy (the marched state keeps drifting ghost values)")
    ap.add_argument("--moving-interior", action="store_true", help="initialise the cells inside the body with the free stream too")
    args = ap.parse_args()
    import immersedboundary_jl_b200 as ib
    import oracle
    from oracle import cfd, cpu_ref
    M = ib
    RAE = os.path.join(ROOT, "tests", "golden", "rae2822.dat")
    stl = M.merge_points(M.Stereolitography(RAE))
    feat = M.DistanceField(M.feature_regions(stl, radius=0.05))
    msh = M.Mesh(np.array([-25, -25], F32), np.array([50, 50], F32), ("wall", stl, F32(1e-2)), refinement_regions=[(feat, F32(5e-3))])
    fams = [("farfield", [(0, False), (0, True), (1, False), (1, True)])]
    dom = ib.Domain(msh, max_partition_size=10_000, hypercube_families=fams, upload=False, ghost_layer_ratio=F32(args.glr))
    ref = cpu_ref.CpuRef.from_builder(dom)
    N = len(dom)
    fl = cfd.Fluid()
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    al = np.radians(args.alpha)
    Pinf = np.array([101325.0, 288.15, args.mach * a_inf * np.cos(al), args.mach * a_inf * np.sin(al)], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", cfd.FlowBC(fl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(fl, Pinf))]
    Pfull = Pinf.copy()
    if args.ramp:
        Pinf = Pfull.copy()
        Pinf[2:] *= F32(0.2)
        bcs = [("wall", cfd.FlowBC(fl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(fl, Pinf))]
    P0 = np.tile(Pinf, (N, 1))
    if not args.moving_interior:
        # cells inside the airfoil form a closed cavity: start them at rest (an impulsively moving gas sloshes there
        # forever and evacuates the leading-edge tip)
        P0[ib.synthetic.inside_polygon(np.loadtxt(RAE), dom.cells()[0]), 2:] = 0
    Q = np.asfortranarray(ib.synthetic.primitive2state_host(P0))
    done = 0
    if args.load:
        prev = np.load(args.load)
        Q, done = np.asfortranarray(prev["Q"]), int(prev["steps"])
    s = dom.surfaces["wall"]
    alphas = {1: [1.0], 2: [0.5, 1.0], 3: [0.1481, 0.4, 1.0], 4: [0.25, 1 / 3, 0.5, 1.0], 5: [0.25, 1 / 6, 0.375, 0.5, 1.0]}[args.stages]
    cflv = F32(args.cfl)
    R, cf = np.zeros((N, 4), F32, order="F"), np.zeros(N, F32)

    def resid(Qs):
        # impose_bc! mutates the state it is given (test/advection.jl:28-46 applies the BCs to `u` itself before marching)
        Qg = Qs.copy(order="F") if args.copy_ghosts else Qs
        ref.ghost_update(fl, Qg, bcs, args.threads)
        ref.residual(fl, Qg, R, cf, args.threads, use_sensor=not args.no_sensor)
        return R, cf

    def coeffs(Qs):
        P = cfd.state2primitive(fl, Qs)
        cp = cfd.pressure_coefficient(fl, P[:, 0], Pinf[0], args.mach)
        cps = np.array([(cp[s.idx[a:b]] * s.w[a:b]).sum() for a, b in zip(s.ptr[:-1], s.ptr[1:])], F32)
        Fxy = (cps[:, None] * s.normals * s.areas[:, None]).sum(axis=0)
        return float(-Fxy[0] * np.sin(al) + Fxy[1] * np.cos(al)), float(Fxy[0] * np.cos(al) + Fxy[1] * np.sin(al))

    ghost = np.zeros(N, bool)
    for chunks in dom.boundaries.values():
        for b in chunks.values():
            ghost[b.ghost_indices] = True
    live = (~ghost if not args.evolve_ghosts else np.ones(N, bool)).astype(F32)[:, None]
    t0 = time.time()
    r0 = None
    hist = []
    for it in range(args.steps + 1):
        if args.ramp:
            fac = F32(min(1.0, 0.2 + 0.8 * it / args.ramp))
            Pnow = Pfull.copy()
            Pnow[2:] *= fac
            bcs[1] = ("farfield", cfd.FlowBC(fl, Pnow))
        if not args.copy_ghosts:
            ref.ghost_update(fl, Q, bcs, args.threads)      # so that Q0 below carries this step's ghost values
        Q0 = Q.copy(order="F")
        for a in alphas:
            Rr, c = resid(Q)
            Q = np.asfortranarray(Q0 + (F32(a) * cflv / c)[:, None] * Rr * live)
        if not np.isfinite(Q).all():
            bad = np.flatnonzero(~np.isfinite(Q).all(axis=1))
            X = dom.cells()[0]
            print(f"step {it}: diverged (non-finite state) at {len(bad)} cells, first at x = {X[bad[:4]].tolist()}, ghost = {ghost[bad[:4]].tolist()}")
            Pp = cfd.state2primitive(fl, Qprev)
            k = int(np.argmin(Pp[:, 0]))
            print(f"  previous step: min p = {Pp[k, 0]:.1f} at x = {X[k].tolist()} (ghost = {bool(ghost[k])}); min rho = {Qprev[:, 0].min():.4f}")
            return
        Qprev = Q
        if it % args.every == 0:
            Qm = Q.copy(order="F")                      # monitor on a copy: a second ghost update would perturb the march
            ref.ghost_update(fl, Qm, bcs, args.threads)
            ref.residual(fl, Qm, R, cf, args.threads, use_sensor=not args.no_sensor)
            Rr, c = R, cf
            nr = float(np.linalg.norm(((Rr / c[:, None]) * live)[:, 0]))
            r0 = r0 or nr
            try:
                cl, cd = coeffs(Q)
            except Exception as e:  # noqa: BLE001
                cl = cd = float("nan")
                if it == 0:
                    print("coeffs unavailable:", e)
            rho = Q[:, 0]
            hist.append((done + it + 1, nr, cl, cd))
            print(f"step {it:6d}  |R_rho/cfl| {nr:.4e}  ratio {nr / r0:.3e}  Cl {cl:+.6f}  Cd {cd:+.6f}  rho[min,max]=({rho.min():.3f},{rho.max():.3f})  t={time.time() - t0:.0f}s", flush=True)
    if args.save:
        cl, cd = coeffs(Q)
        h = np.array(hist)
        last = h[h[:, 0] > (done + args.steps + 1) * 2 // 3]
        np.savez_compressed(args.save, Q=Q, cl=cl, cd=cd, steps=done + args.steps + 1, history=h,
                            cl_drift=float(np.ptp(last[:, 2])), cd_drift=float(np.ptp(last[:, 3])), command=" ".join(sys.argv), stages=args.stages, cfl=args.cfl, mach=args.mach, alpha=args.alpha)


if __name__ == "__main__":
    main()
