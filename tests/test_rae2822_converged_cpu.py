"""The C3 steady-state fixture (tests/golden/rae2822_converged.npz, written by tools/c3_converge.py) is a fixed point of the
oracle's march: Cl / Cd recomputed from the stored state are the stored ones, and 60 more steps of the same driver leave
them within the drift the fixture records.  CPU only (compiled restatement oracle/cpu_ref.c on the product's host tables)."""
import os

import numpy as np

F32 = np.float32
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rae2822_converged.npz")


def test_fixture_is_a_steady_state_of_the_oracle_march(get_case, ib, oracle):
    from oracle import cfd, cpu_ref
    g = np.load(GOLDEN)
    dom = get_case("rae2822", 10_000).dom
    N = len(dom)
    assert g["Q"].shape == (N, 4) and int(g["steps"]) == 120_000
    mach, alpha, CFL = float(g["mach"]), float(g["alpha"]), F32(g["cfl"])
    fl = cfd.Fluid()
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    al = np.radians(alpha)
    Pinf = np.array([101325.0, 288.15, mach * a_inf * np.cos(al), mach * a_inf * np.sin(al)], F32)
    bcs = [("wall", cfd.FlowBC(fl, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)), ("farfield", cfd.FlowBC(fl, Pinf))]
    s = dom.surfaces["wall"]

    def coeffs(Q):
        cp = cfd.pressure_coefficient(fl, cfd.state2primitive(fl, Q)[:, 0], Pinf[0], mach)
        cps = np.array([(cp[s.idx[a:b]] * s.w[a:b]).sum() for a, b in zip(s.ptr[:-1], s.ptr[1:])], F32)
        F = (cps[:, None] * s.normals * s.areas[:, None]).sum(axis=0)
        return float(-F[0] * np.sin(al) + F[1] * np.cos(al)), float(F[0] * np.cos(al) + F[1] * np.sin(al))

    Q = np.asfortranarray(g["Q"])
    cl, cd = coeffs(Q)
    assert abs(cl - float(g["cl"])) < 1e-6 and abs(cd - float(g["cd"])) < 1e-6
    assert 0.85 < -cl < 0.92 and 0.02 < -cd < 0.03          # transonic RAE2822, inviscid: lift ~0.89, wave drag ~0.023
    ghost = np.zeros(N, bool)
    for chunks in dom.boundaries.values():
        for b in chunks.values():
            ghost[b.ghost_indices] = True
    live = (~ghost).astype(F32)[:, None]
    ref = cpu_ref.CpuRef.from_builder(dom)
    R, cf = np.zeros((N, 4), F32, order="F"), np.zeros(N, F32)
    for _ in range(60):
        ref.ghost_update(fl, Q, bcs)
        Qs = Q.copy(order="F")
        for a in ib.RK_STAGES[int(g["stages"])]:
            ref.ghost_update(fl, Q, bcs)
            ref.residual(fl, Q, R, cf)
            Q = np.asfortranarray(Qs + (F32(a) * CFL / cf)[:, None] * R * live)
    cl2, cd2 = coeffs(Q)
    assert abs(cl2 - cl) < float(g["cl_drift"]) + 1e-5 and abs(cd2 - cd) < float(g["cd_drift"]) + 1e-5, (cl, cl2, cd, cd2)
    # the history shows the approach: lift within 1e-4 of the final value from 50 000 steps on
    h = g["history"]
    assert np.abs(h[h[:, 0] > 50_000, 2] - cl).max() < 1e-4
