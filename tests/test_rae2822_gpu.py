"""C3 (BASELINE.json configs[2]): transonic RAE2822 Euler with immersed-boundary ghost cells, marched with local time
steps through the drop-in boundary (`ibx_euler_step_host`: host state in, residual + CFL term out) next to the oracle
doing the same march, then lift / drag from the surface pressure integral (Surface + pressure_coefficient +
surface_integral).  Tolerances asserted below: state within 5e-4 of its per-variable scale, lift and drag coefficients
within 1e-4 (the north-star figure for Cl / Cd).

The reference ships no Euler residual and no time integrator.  This file keeps the round-1 comparison on
the start-up transient of the plain forward-Euler march in which EVERY cell -- ghost cells included -- is advanced by
its residual; that driver loses its ghost-cell values next to the thin trailing edge after ~28 steps in the ORACLE as
well (a ghost cell advanced by its own residual drifts, and image points interpolate from it).
The converged polar is in tests/test_rae2822_converged_gpu.py (ghost cells frozen between residual evaluations).  Here
both paths have processed the same 20 ghost updates + residuals; the 2e-6 difference of the ghost-interpolation weights
is amplified to ~1e-4 of the state scale by then."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32
MACH, ALPHA, CFL, STEPS = 0.73, 2.31, F32(0.4), 20


def _coefficients(Fxy):
    a = np.radians(ALPHA)
    return float(-Fxy[0] * np.sin(a) + Fxy[1] * np.cos(a)), float(Fxy[0] * np.cos(a) + Fxy[1] * np.sin(a))


def test_rae2822_march_lift_and_drag(get_case, ib, oracle):
    c = get_case("rae2822", 10_000, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    N = len(c.dom)
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    d = ib.streamwise_direction(ALPHA)
    Pinf = np.array([101325.0, 288.15, MACH * a_inf * d[0], MACH * a_inf * d[1]], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    obcs = [("wall", cfd.FlowBC(ofl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(ofl, Pinf))]
    Q0 = np.asfortranarray(ib.synthetic.primitive2state_host(np.tile(Pinf, (N, 1))))
    # ---- product: every step crosses the C ABI with host buffers
    Q = ib.pinned_empty((N, 4))
    Q[...] = Q0
    R, cf = ib.pinned_empty((N, 4)), ib.pinned_empty((N,))
    for _ in range(STEPS):
        ib.euler_step_host(c.dom, fl, bcs, Q, R, cf)
        Q += (CFL / cf)[:, None] * R
    # ---- oracle: the same march (ghost update on a copy, like the device-side staging array)
    Qo = Q0.copy()
    for _ in range(STEPS):
        Qg = Qo.copy()
        E.euler_ghost_update(c.odom, ofl, Qg, obcs)
        Ro, co = np.zeros_like(Qo), np.zeros(N, F32)
        c.odom(E.euler_residual(ofl), Qg, Ro, co)
        Qo += (CFL / co)[:, None] * Ro
    assert np.isfinite(Q).all() and np.abs(Q - Q0).max() > 0
    qs = np.abs(Qo).max(axis=0)
    assert (np.abs(Q - Qo) / qs).max() < 5e-4, (np.abs(Q - Qo) / qs).max()
    # ---- lift and drag from the wall pressure
    s, os_ = c.dom.surfaces["wall"], c.odom.surfaces["wall"]
    p = ib.state2primitive(fl, ib.DeviceArray.from_host(np.asfortranarray(Q))).col(0)
    Cp_s = s(ib.pressure_coefficient(fl, p, Pinf[0], MACH)).to_host().ravel()
    F = ib.surface_integral(s, np.asfortranarray(Cp_s[:, None] * s.normals))
    po = cfd.state2primitive(ofl, Qo)[:, 0]
    Cpo_s = os_(cfd.pressure_coefficient(ofl, po, Pinf[0], MACH))
    Fo = oracle.domain.surface_integral(os_, Cpo_s[:, None] * os_.normals)
    (cl, cd), (clo, cdo) = _coefficients(F), _coefficients(Fo)
    assert abs(cl - clo) < 1e-4 and abs(cd - cdo) < 1e-4, (cl, clo, cd, cdo)
    assert abs(cl) > 1e-4                                        # the incidence produces lift within the first steps


def test_rae2822_fas_multigrid_cycle(get_case, ib, oracle):
    """C3 "with mgrid multigrid": one FAS V-cycle (src/solver.jl:39-91) over the block_size 8 -> 4 -> 2 hierarchy of
    multigrid(dom) (src/ImmersedBoundary.jl:1355-1407) with the Euler residual + IB ghost update on every level,
    entirely on device arrays (restriction / prolongation through the IDW transfer accumulators, clamped update,
    norms), next to the oracle running the same cycle."""
    c = get_case("rae2822", 10_000, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    d = ib.streamwise_direction(ALPHA)
    Pinf = np.array([101325.0, 288.15, MACH * a_inf * d[0], MACH * a_inf * d[1]], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    obcs = [("wall", cfd.FlowBC(ofl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(ofl, Pinf))]
    cd, pro, coa = ib.multigrid(c.dom)
    ocd, opro, ocoa = oracle.domain.multigrid(c.odom)
    doms, odoms = [c.dom] + cd, [c.odom] + ocd
    assert [len(x) for x in doms] == [len(x.centers) for x in odoms] and len(doms) == 4
    doms, odoms, coa, pro, ocoa, opro = doms[:3], odoms[:3], coa[:2], pro[:2], ocoa[:2], opro[:2]

    def f(l, Q):
        Qg = Q.copy()
        ib.ghost_update_euler(doms[l], fl, Qg, bcs)
        R, cf = ib.DeviceArray(Q.rows, 4, False), ib.DeviceArray(Q.rows, 1, True)
        ib.residual_euler(doms[l], fl, Qg, R, cf)
        return R * (float(CFL) / cf), 1.0

    def fo(l, Q):
        Qg = Q.copy()
        E.euler_ghost_update(odoms[l], ofl, Qg, obcs)
        R, cf = np.zeros_like(Q), np.zeros(len(Q), F32)
        odoms[l](E.euler_residual(ofl), Qg, R, cf)
        return R * (CFL / cf)[:, None], F32(1.0)

    Q0 = ib.synthetic.primitive2state_host(np.tile(Pinf, (len(c.dom), 1)))
    Q = ib.DeviceArray.from_host(Q0)
    ratio = ib.FAS(f, Q, coa, pro, n_iter=3, rtol=1e-6)
    Qo = Q0.copy()
    oratio = oracle.solver.FAS(fo, Qo, ocoa, opro, n_iter=3, rtol=F32(1e-6))
    Qh = Q.to_host()
    qs = np.abs(Qo).max(axis=0)
    assert np.isfinite(Qh).all() and np.abs(Qh - Q0).max() > 0
    # the transfer and ghost-interpolation weights agree to ~1e-6 (float32 SVD vs double Jacobi); a residual amplifies a
    # relative change of its input by ~1e2, and the cycle chains ~12 evaluations over three levels
    assert (np.abs(Qh - Qo) / qs).max() < 5e-4, (np.abs(Qh - Qo) / qs).max()
    assert abs(float(ratio) - float(oratio)) < 1e-3 * max(float(oratio), 1e-6), (ratio, oratio)


def test_rae2822_export_vtk_like_the_reference_script(get_case, ib, tmp_path):
    """The closing lines of test/rae2822.jl: `export_vtk("rae2822", dom; ny = ny)` with ny from impose_bc!, volume and
    surface files, and the coarsest multigrid domain."""
    import os
    import xml.etree.ElementTree as ET
    c = get_case("rae2822", 10_000, upload=True)
    N = len(c.dom)
    ny = np.zeros(N, F32)
    ib.impose_bc(lambda b, x: b.normals.col(1), c.dom, "wall", ny)          # test/rae2822.jl:31-34
    out = ib.export_vtk(str(tmp_path / "rae2822"), c.dom, ny=ny, surface_data={"wall": {"area": c.dom.surfaces["wall"].areas}})
    assert len(ET.parse(out["volume"]).getroot().findall(".//DataSet")) == c.msh.nblocks == 580
    surf = ET.parse(out["surface"]).getroot().findall(".//DataSet")
    assert [s.get("name") for s in surf] == ["wall"]
    piece = ET.parse(os.path.join(str(tmp_path / "rae2822"), surf[0].get("file"))).getroot().find(".//Piece")
    named = {a.get("Name"): np.array(a.text.split(), dtype=np.float64) for tag in ("PointData", "CellData") for a in piece.find(tag)}
    s = c.dom.surfaces["wall"]
    assert np.allclose(named["ny"], s(ny), atol=1e-7) and np.allclose(named["area"], s.areas)
    assert np.abs(named["ny"]).max() > 0.5                                  # the wall normals reach the surface values
    cd, _, _ = ib.multigrid(c.dom)
    out = ib.export_vtk(str(tmp_path / "rae2822_coarse"), cd[-1], export_surface=False)
    assert len(ET.parse(out["volume"]).getroot().findall(".//DataSet")) == 580


def test_2d_residual_is_independent_of_earlier_calls(get_case, ib):
    c2 = get_case("rae2822", 10_000, upload=True)
    c3 = get_case("sphere3d", 40_000, upload=True)
    fl = ib.Fluid()

    def run2():
        N = len(c2.dom)
        Q = ib.DeviceArray.from_host(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c2.dom.cells()[0])))
        R, cf = ib.DeviceArray(N, 4, False), ib.DeviceArray(N, 1, True)
        ib.residual_euler(c2.dom, fl, Q, R, cf)
        return R.to_host(), cf.to_host()

    a = run2()
    junk = [ib.DeviceArray(3_000_000, 5, False).fill(float("nan")) for _ in range(3)]   # poison freed memory
    N3 = len(c3.dom)
    Q3 = ib.DeviceArray.from_host(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c3.dom.cells()[0])) * F32(1.3))
    R3, c3f = ib.DeviceArray(N3, 5, False), ib.DeviceArray(N3, 1, True)
    ib.residual_euler(c3.dom, fl, Q3, R3, c3f)                                           # different scratch layout
    del junk
    b = run2()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (int((a[0] != b[0]).sum()), float(np.nanmax(np.abs(a[0] - b[0]))))


def test_point_implicit_on_the_euler_residual_vs_oracle(get_case, ib, oracle):
    """hutchinson_trick / linearize / proj_along / solve (src/point_implicit.jl:18-91, 184-233, 250-329) driven by the Euler
    residual + IB ghost update of C3 on the device, next to oracle/point_implicit.py doing the same with the same
    counter-based +-1 probes (synthetic.probe_signs, keyed by the global cell id).  Unknowns are scaled to O(1) so that the
    finite-difference step is resolvable in Float32 (the reference's Float64 `h = 1e-6` would promote the evaluation)."""
    from oracle import point_implicit as opi
    c = get_case("rae2822", 10_000, upload=True)
    E, cfd = oracle.euler, oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    N, nv = len(c.dom), 4
    a_inf = np.sqrt(1.4 * 283.0 * 288.15)
    d = ib.streamwise_direction(ALPHA)
    Pinf = np.array([101325.0, 288.15, MACH * a_inf * d[0], MACH * a_inf * d[1]], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    obcs = [("wall", cfd.FlowBC(ofl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(ofl, Pinf))]
    P0 = ib.synthetic.euler_state(c.odom.centers, mach=MACH)
    Q0 = ib.synthetic.primitive2state_host(P0)
    scale = np.abs(Q0).max(axis=0).astype(F32)
    X0 = (Q0 / scale).astype(F32)
    S = ib.DeviceArray.from_host(np.tile(scale, (N, 1)))
    h = F32(2e-3)

    def f(X):                                   # pseudo-time increment of the scaled unknowns
        Qg = X * S
        ib.ghost_update_euler(c.dom, fl, Qg, bcs)
        R, cf = ib.DeviceArray(N, nv, False), ib.DeviceArray(N, 1, True)
        ib.residual_euler(c.dom, fl, Qg, R, cf)
        return R * (float(CFL) / cf) / S

    def fo(X):
        Qg = (X * scale).astype(F32)
        E.euler_ghost_update(c.odom, ofl, Qg, obcs)
        R, cf = np.zeros_like(Qg), np.zeros(N, F32)
        c.odom(E.euler_residual(ofl), Qg, R, cf)
        return (R * (CFL / cf)[:, None] / scale).astype(F32)

    # Ghost-interpolation weights agree to ~2e-6 only (float32 SVD in the oracle, double Jacobi in the product) and the
    # residual amplifies that by ~1e3 next to the body: give the oracle the product's weights (donor sets are identical,
    # tests/test_builder_parity.py) so that the comparison below is about the solver chain, not about pinv roundings.
    saved = {}
    for bname, chunks in c.odom.boundaries.items():
        for k, ob in chunks.items():
            b = c.dom.boundaries[bname][k]
            saved[(bname, k)] = ob.image_interpolator
            ob.image_interpolator = oracle.accumulator.Accumulator.from_csr(b.interp_ptr, b.interp_idx, b.interp_w)
    try:
        _point_implicit_body(ib, oracle, opi, f, fo, X0, N, nv, h)
    finally:
        for (bname, k), acc in saved.items():
            c.odom.boundaries[bname][k].image_interpolator = acc


def _point_implicit_body(ib, oracle, opi, f, fo, X0, N, nv, h):
    n_samples = 2
    probes = ib.synthetic.probe_signs(np.arange(N), nv, n_samples, seed=3)
    X = ib.DeviceArray.from_host(X0)
    fX, foX = f(X), fo(X0)
    fs = np.abs(foX).max()
    assert np.abs(fX.to_host() - foX).max() < 1e-5 * fs
    D = ib.hutchinson_trick(f, X, n_samples, h=h, fX=fX, probes=probes).to_host().reshape(N, nv, nv).transpose(0, 2, 1)
    Do = opi.hutchinson_trick(fo, X0, n_samples, h, foX, probes=probes)        # D[p, j, i]
    ds = np.abs(Do).max()
    # finite differences divide the Float32 rounding of f by h: agreement to ~1e-7 |f| / h
    assert np.abs(D - Do).max() < 2e-2 * ds, (np.abs(D - Do).max(), ds)
    assert np.median(np.abs(D - Do)) < 1e-4 * ds
    lin, b, pre = ib.linearize(f, X, n_hutchinson_samples=n_samples, pre_evaluated_fx=fX, h=h, probes=probes)
    olin, ob, opre = opi.linearize(fo, X0, n_samples, foX, h, probes=probes)
    v = (probes[0].T * F32(0.01)).astype(F32)
    Av, oAv = lin(ib.DeviceArray.from_host(v)).to_host(), olin(v)
    assert np.abs(Av - oAv).max() < 2e-2 * np.abs(oAv).max()
    dx, ratio = ib.solve(lin, b, pre, n_iter=3, rtol=1e-6)
    odx, oratio = opi.solve(olin, ob, opre, n_iter=3, rtol=F32(1e-6))
    assert np.isfinite(dx.to_host()).all()
    assert abs(float(ratio) - float(oratio)) < 5e-2 * max(float(oratio), 1e-3), (ratio, oratio)
    assert float(ratio) < 1.0                                     # the preconditioned projection steps reduce the residual
