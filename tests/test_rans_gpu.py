"""Configuration C5 (BASELINE.json configs[4]): the canonical RANS residual -- Euler part + viscous fluxes with the
Wray-Agarwal eddy viscosity + the transported-R residual -- on the device (ibx_residual_rans: fused Euler kernels +
table-free viscous / turbulence kernels, rans.cu) against the oracle composition of the restated reference operators
(oracle/euler.py: rans_residual).  A deliberately viscous fluid (mu_ref 2e-2) makes the viscous terms as large as the
inviscid ones, so that the tolerance means something for them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.mark.parametrize("name,mps", [("sphere3d_stl", 100_000), ("sphere3d", 40_000)])
def test_rans_residual_against_oracle(get_case, ib, oracle, name, mps):
    c = get_case(name, mps, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl = ib.Fluid(mu_ref=2e-2)
    ofl = cfd.Fluid(mu_ref=F32(2e-2))
    od = c.odom
    N = len(od.centers)
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(od.centers))
    x = od.centers
    qR0 = (Q0[:, 0] * F32(3e-2) * (1 + F32(0.3) * np.sin(F32(2) * x[:, 0]) * np.cos(F32(3) * x[:, 1]))).astype(F32)
    Ro, RRo, co = np.zeros_like(Q0), np.zeros(N, F32), np.zeros(N, F32)
    od(E.rans_residual(ofl), Q0.copy(), qR0.copy(), Ro, RRo, co)
    Re, ce = np.zeros_like(Q0), np.zeros(N, F32)
    od(E.euler_residual(ofl), Q0.copy(), Re, ce)
    Q, qR = ib.DeviceArray.from_host(Q0), ib.DeviceArray.from_host(qR0)
    R, RR, cf = ib.DeviceArray(N, 5, False), ib.DeviceArray(N, 1, True), ib.DeviceArray(N, 1, True)
    ib.residual_rans(c.dom, fl, Q, qR, R, RR, cf)
    Rg, RRg, cg = R.to_host(), RR.to_host(), cf.to_host()
    assert np.array_equal(cg, co)
    # the viscous increment on its own (the Euler part is bit-exact, tests/test_fused_gpu.py)
    dv, dvo = Rg - Re, Ro - Re
    vs = np.abs(dvo).max(axis=0)
    assert (vs[1:] > 1e-4 * np.abs(Re).max(axis=0)[1:]).all()                 # the viscous terms are not negligible here
    assert np.array_equal(Rg[:, 0], Ro[:, 0])                                   # no viscous mass flux
    # R = ((R_euler + gg_1) + gg_2) + gg_3 in Float32: each sum rounds to an ulp of R, which is larger than 1e-5 of the
    # viscous increment where the inviscid residual dominates -- allow those roundings on top of the tolerance
    tol = 1e-5 * vs[1:] + 3 * np.spacing(np.abs(Ro[:, 1:]))
    excess = np.abs(Rg[:, 1:] - Ro[:, 1:]) - tol
    assert excess.max() <= 0, (excess.max(axis=0), vs)
    assert np.median(np.abs(dv[:, 1:] - dvo[:, 1:]) / vs[1:]) < 1e-6
    rs = np.abs(RRo).max()
    assert rs > 0 and np.abs(RRg - RRo).max() < 1e-5 * rs, (np.abs(RRg - RRo).max(), rs)
    # whole residual under the north-star tolerance (flux-scaled)
    from bench import flux_scaled_error
    assert flux_scaled_error(od.widths, Q0, Rg, Ro).max() < 1e-5


def test_rans_ghost_update_against_oracle(get_case, ib, oracle):
    c = get_case("sphere3d", 40_000, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    od = c.odom
    N = len(od.centers)
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(od.centers))
    qR0 = (Q0[:, 0] * F32(3e-4) * (1 + F32(0.3) * np.sin(F32(2) * od.centers[:, 0]))).astype(F32)
    a = np.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    obcs = [("wall", cfd.FlowBC(ofl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(ofl, Pinf))]
    rbc = [("wall", 0.0), ("farfield", 4.5e-5)]
    Q, qR = ib.DeviceArray.from_host(Q0), ib.DeviceArray.from_host(qR0)
    ib.ghost_update_euler(c.dom, fl, Q, bcs)
    ib.ghost_update_rans(c.dom, Q, qR, rbc)
    Qo, qRo = Q.to_host(), qR0.copy()          # the oracle continues from the SAME ghost-updated mean-flow state
    E.rans_ghost_update(od, Qo, qRo, rbc)
    got = qR.to_host()
    changed = np.flatnonzero(qRo != qR0)
    assert len(changed) > 1000 and np.array_equal(np.flatnonzero(got != qR0), changed)
    assert np.abs(got - qRo).max() < 2e-6 * np.abs(qRo).max()      # interpolation weights: float32 SVD vs double Jacobi


def test_point_implicit_on_the_rans_residual_vs_oracle(get_case, ib, oracle):
    """C5 "with point_implicit smoothing": linearize (Hutchinson block diagonal, nv = 6: mean flow + transported R) and two
    preconditioned projection sweeps of `solve` (src/point_implicit.jl:184-329) on the RANS residual, device next to
    oracle/point_implicit.py with the same counter-based probes.  Unknowns scaled to O(1) (Float32 finite differences)."""
    from oracle import point_implicit as opi
    c = get_case("sphere3d_stl", 100_000, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl, ofl = ib.Fluid(mu_ref=2e-2), cfd.Fluid(mu_ref=F32(2e-2))
    od = c.odom
    N, nv = len(od.centers), 6
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(od.centers))
    qR0 = (Q0[:, 0] * F32(3e-2) * (1 + F32(0.3) * np.sin(F32(2) * od.centers[:, 0]))).astype(F32)
    X6 = np.concatenate([Q0, qR0[:, None]], axis=1)
    scale = np.abs(X6).max(axis=0).astype(F32)
    X0 = (X6 / scale).astype(F32)
    S5, S1 = ib.DeviceArray.from_host(np.tile(scale[:5], (N, 1))), float(scale[5])
    S6 = ib.DeviceArray.from_host(np.tile(scale, (N, 1)))
    CFL, h = F32(0.4), F32(2e-3)
    res = E.rans_residual(ofl)

    def f(X):
        Xh = X * S6
        Q, qR = ib.DeviceArray(N, 5, False), Xh.col(5)
        for j in range(5):
            cj = Xh.col(j)
            ib._lib.call("ibx_array_set_column", ib.context(), Q.h, j, cj.h)
        R, RR, cf = ib.DeviceArray(N, 5, False), ib.DeviceArray(N, 1, True), ib.DeviceArray(N, 1, True)
        ib.residual_rans(c.dom, fl, Q, qR, R, RR, cf)
        out = ib.DeviceArray(N, 6, False)
        dt = float(CFL) / cf
        Rs = R * dt / S5
        for j in range(5):
            cj = Rs.col(j)
            ib._lib.call("ibx_array_set_column", ib.context(), out.h, j, cj.h)
        r6 = RR * dt / S1
        ib._lib.call("ibx_array_set_column", ib.context(), out.h, 5, r6.h)
        return out

    def fo(X):
        Xh = (X * scale).astype(F32)
        R, RR, cf = np.zeros((N, 5), F32), np.zeros(N, F32), np.zeros(N, F32)
        od(res, Xh[:, :5].copy(), Xh[:, 5].copy(), R, RR, cf)
        dt = CFL / cf
        return np.concatenate([R * dt[:, None] / scale[:5], (RR * dt / scale[5])[:, None]], axis=1).astype(F32)

    probes = ib.synthetic.probe_signs(np.arange(N), nv, 1, seed=5)
    X = ib.DeviceArray.from_host(X0)
    fX, foX = f(X), fo(X0)
    fs = np.abs(foX).max(axis=0)
    assert (np.abs(fX.to_host() - foX) / fs).max() < 2e-5
    lin, b, pre = ib.linearize(f, X, n_hutchinson_samples=1, pre_evaluated_fx=fX, h=h, probes=probes)
    olin, ob, opre = opi.linearize(fo, X0, 1, foX, h, probes=probes)
    v = (probes[0].T * F32(0.01)).astype(F32)
    Av, oAv = lin(ib.DeviceArray.from_host(v)).to_host(), olin(v)
    assert np.abs(Av - oAv).max() < 2e-2 * np.abs(oAv).max()
    dx, ratio = ib.solve(lin, b, pre, n_iter=2, rtol=1e-6)
    odx, oratio = opi.solve(olin, ob, opre, n_iter=2, rtol=F32(1e-6))
    assert np.isfinite(dx.to_host()).all() and float(ratio) < 1.0
    assert abs(float(ratio) - float(oratio)) < 5e-2 * max(float(oratio), 1e-3), (ratio, oratio)
