"""The oracle against the reference's only known-answer vector and the analytic invariants of SURVEY.md 8(c)."""
import json
import os

import numpy as np
import pytest

F32 = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))


def test_accumulator_kat(oracle):
    kat = json.load(open(os.path.join(HERE, "golden", "accumulator_kat.json")))
    st = [[i - 1 for i in s] for s in kat["stencils_1based"]]
    acc = oracle.accumulator.Accumulator(st, kat["weights"])
    assert np.array_equal(acc(np.array(kat["v"])), np.array(kat["expected"]))  # src/accumulator.jl:25-34


def test_accumulator_delta_and_unweighted(oracle):
    acc = oracle.accumulator.Accumulator([[0, 1], [1, 2, 3]], [[1.0, 1.0], [1.0, 1.0, 1.0]])
    v = np.array([1.0, 2.0, 4.0, 8.0])
    assert np.array_equal(acc(v, delta=True), np.array([(1 - 1) + (2 - 1), (2 - 2) + (4 - 2) + (8 - 2)], float))
    un = oracle.accumulator.Accumulator([[0, 1], [1, 2, 3]])
    assert np.array_equal(un(v), np.array([3.0, 14.0]))
    m = np.stack([v, 2 * v], axis=1)
    acc_f = oracle.accumulator.Accumulator([[0, 1], [1, 2, 3]], [[1.0, 1.0], [1.0, 1.0, 1.0]], first_index=True)
    assert np.array_equal(acc_f(m)[:, 1], 2 * acc_f(m)[:, 0])


def test_mesh_sizes_match_survey(get_case):
    # advection 187 blocks / 11 968 cells, rae2822 580 blocks / 37 120 cells over 10 levels
    c = get_case("advection")
    assert c.omsh.block_origins.shape[0] == 187 and len(c.omsh) == 11968
    r = get_case("rae2822")
    assert r.omsh.block_origins.shape[0] == 580 and len(r.omsh) == 37120
    assert len(np.unique(r.omsh.block_widths[:, 0])) == 10


def test_volume_integral_and_centroid(get_case, oracle):
    r = get_case("rae2822")
    D = oracle.domain
    V = D.volume_integral(r.odom, np.ones(len(r.odom), F32))
    assert abs(V - 2500.0) < 2500 * 1e-5
    CG = D.volume_integral(r.odom, r.odom.centers.copy()) / F32(2500)  # test/rae2822.jl:24-29, analytic (0, 0)
    assert np.all(np.abs(CG) < 1e-3)


def test_gradient_of_linear_field_and_constant(get_case, oracle):
    c = get_case("advection")
    D = oracle.domain
    part = c.odom.partitions[1]
    lin = (F32(2) * part.centers[:, 0] + F32(3) * part.centers[:, 1]).astype(F32)
    g0, g1 = D.cell_gradient(part, lin, 0), D.cell_gradient(part, lin, 1)
    # exact in uniform regions (cells whose 4 neighbours have the same width)
    uniform = np.ones(len(lin), bool)
    for dim in range(2):
        o, n = part.face_owners_neighbors[dim]
        bad = part.spacing[o, dim] != part.spacing[n, dim]
        uniform[o[bad]] = False
        uniform[n[bad]] = False
        box = o == n
        uniform[o[box]] = False
    assert np.allclose(g0[uniform], 2, atol=2e-4) and np.allclose(g1[uniform], 3, atol=2e-4)
    const = np.full(len(lin), F32(7.5))
    assert np.abs(D.cell_gradient(part, const, 0)).max() < 1e-3


def test_interpolator_weights(get_case, oracle):
    r = get_case("rae2822")
    b = r.odom.boundaries["wall"][1]
    ptr, idx, w = b.image_interpolator.to_csr()
    sums = np.add.reduceat(w, ptr[:-1])
    assert np.allclose(sums, 1.0, atol=2e-5)  # linear weights sum to one
    # and reproduce linear fields at the image points
    X = r.odom.centers[b.image_domain]
    f = (F32(0.3) * X[:, 0] - F32(1.7) * X[:, 1] + F32(2)).astype(F32)
    fi = b.image_interpolator(f)
    exact = 0.3 * b.images[:, 0] - 1.7 * b.images[:, 1] + 2
    assert np.allclose(fi, exact, atol=5e-4)


def test_impose_bc_invariants(get_case, oracle):
    c = get_case("advection")
    D = oracle.domain
    rng = np.random.default_rng(0)
    u = rng.random(len(c.odom)).astype(F32)
    b = c.odom.boundaries["outlet"][1]
    ia = b.image_interpolator(u[b.image_domain])
    u2 = u.copy()
    D.impose_bc(lambda bd, ui: ui.copy(), c.odom, "outlet", u2)  # f = copy  =>  ghost = image value
    assert np.allclose(u2[b.ghost_indices], ia, atol=1e-6)
    u3 = u.copy()
    D.impose_bc(lambda bd, ui: F32(1.0), c.odom, "outlet", u3)   # scalar  =>  eta ia + (1 - eta) c
    eta = b.ghost_distances / b.image_distances
    assert np.allclose(u3[b.ghost_indices], eta * ia + (1 - eta) * 1.0, atol=1e-6)
    assert np.all((eta >= 0) & (eta <= 1.0 + 1e-6))


def test_hll_consistency_and_muscl_linear(get_case, oracle):
    cfd = oracle.cfd
    fl = cfd.Fluid()
    P = np.array([[101325.0, 288.15, 120.0, -30.0], [90000.0, 250.0, 400.0, 10.0]], F32)
    for dim in range(2):
        F = cfd.inviscid_fluxes_hll(fl, P, P, dim)
        Q = cfd.primitive2state(fl, P)
        phys = Q.astype(np.float64).copy()
        phys[:, 1] += P[:, 0]
        phys *= P[:, 2 + dim][:, None]
        phys[:, 2 + dim] += P[:, 0]
        assert np.allclose(F, phys, rtol=1e-6)  # PL == PR  =>  physical flux
    assert F.dtype == np.float64                # the reference's Float64 promotion (src/cfd.jl:504-507)
    assert np.allclose(cfd.state2primitive(fl, cfd.primitive2state(fl, P)), P, rtol=1e-6)
    # MUSCL of a linear field with exact gradient returns the face value on uniform faces
    c = get_case("advection")
    D = oracle.domain
    part = c.odom.partitions[1]
    a = F32(1.5)
    u = (a * part.centers[:, 0]).astype(F32)
    du = np.full(len(u), a, F32)
    uL, uR = D.MUSCL(part, u, du, 0)
    o, n = part.face_owners_neighbors[0]
    uni = (part.spacing[o, 0] == part.spacing[n, 0]) & (o != n)
    uf = D.at_faces(part, u, 0)
    assert np.allclose(uL[uni], uf[uni], atol=1e-5) and np.allclose(uR[uni], uf[uni], atol=1e-5)


def test_partition_independence_of_residual(get_case, oracle):
    """Image-cell results do not depend on the partitioning (2-deep skirt) -- the property the fused path uses."""
    one, many = get_case("rae2822"), get_case("rae2822", mps=10_000)
    E, cfd = oracle.euler, oracle.cfd
    fl = cfd.Fluid()
    from immersedboundary_jl_b200 import synthetic
    Q = synthetic.primitive2state_host(synthetic.euler_state(one.odom.centers))
    out = []
    for case in (one, many):
        R, cf = np.zeros_like(Q), np.zeros(len(Q), F32)
        case.odom(E.euler_residual(fl), Q.copy(), R, cf)
        out.append((R, cf))
    scale = np.abs(out[0][0]).max(axis=0)
    assert (np.abs(out[0][0] - out[1][0]) / scale).max() < 1e-6
    assert np.allclose(out[0][1], out[1][1], rtol=1e-6)


def test_multigrid_restriction_weights(get_case, oracle):
    r = get_case("advection")
    cd, pro, coa = oracle.domain.multigrid(r.odom)       # code order: (coarse_doms, prolongators, coarseners)
    assert [len(c) for c in cd] == [len(r.odom) // 4, len(r.odom) // 16, len(r.odom) // 64]
    ptr, idx, w = coa[0].to_csr()
    assert np.all(np.diff(ptr) == 4) and np.allclose(w, 0.25, atol=1e-6)  # 2^nd equidistant children
    one = np.ones(len(r.odom), F32)
    assert np.allclose(pro[0](coa[0](one)), 1.0, atol=1e-5)


def test_fas_and_point_implicit_on_linear_problem(oracle):
    rng = np.random.default_rng(1)
    n, nv = 40, 3
    A = rng.random((n, nv, nv)).astype(F32) * F32(0.1) + np.eye(nv, dtype=F32)[None] * F32(2)
    xs = rng.random((n, nv)).astype(F32)
    f = lambda x: np.einsum("pji,pi->pj", A, x - xs).astype(F32)
    pi = oracle.point_implicit
    x0 = np.zeros((n, nv), F32)
    D = pi.hutchinson_trick(f, x0, 8, h=1e-3)
    assert np.allclose(D, A, atol=2e-2)                       # D[p, j, i] ~ d f_j / d x_i
    lin, b, prec = pi.linearize(f, x0, n_hutchinson_samples=8, h=1e-3)
    dx, ratio = pi.solve(lin, b, prec, n_iter=50, rtol=1e-3)
    assert ratio < 1e-2 and np.allclose(dx, xs, atol=5e-2)
    Q = np.zeros((n, nv), F32)
    ratio = oracle.solver.FAS(lambda l, Q: (-f(Q), F32(0.4)), Q, n_iter=200, rtol=F32(1e-3))
    assert ratio < 2e-3 and np.allclose(Q, xs, atol=1e-2)


def test_geometric_multigrid_mgrid(oracle):
    rng = np.random.default_rng(2)
    X = rng.random((256, 2)).astype(F32)
    mg = oracle.mgrid.Multigrid(X, 2)
    assert mg.coarseners[0].n_output == 64 and mg.coarseners[1].n_output == 16
    one = np.ones(256, F32)
    assert np.allclose(mg.coarseners[0](one), 1.0, atol=1e-6)
    assert np.allclose(mg.prolongators[0](mg.coarseners[0](one)), 1.0, atol=1e-6)


def test_float64_flux_promotion_is_what_the_tolerance_rests_on(get_case, ib, oracle):
    """DESIGN.md 4.1, on the CPU: the reference's HLL line promotes the flux (and the Green-Gauss sums taken of it) to
    Float64 (src/cfd.jl:504-507).  Rounding that flux to Float32 before the divergence -- the 'obvious' fast kernel --
    moves the residual by more than the north-star tolerance (1e-5 relative per cell) in practically EVERY cell, because
    a residual is a difference of fluxes ~1e3 times larger than itself.  Hence the kernels keep the promotion."""
    from oracle import cfd
    from oracle.domain import JST_sensor, MUSCL, cell_gradient, green_gauss
    c = get_case("sphere3d_stl", 20_000)
    fl = cfd.Fluid()
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c.odom.centers))

    def residual(round_flux):
        def f(part, Q, R):
            P = cfd.state2primitive(fl, Q)
            D = JST_sensor(part, P[:, 0])
            R[...] = 0
            for dim in range(part.ndims):
                PL, PR = MUSCL(part, P, cell_gradient(part, P, dim), dim, D=D, high_order=False)
                Fx = cfd.inviscid_fluxes_hll(fl, PL, PR, dim)
                assert Fx.dtype == np.float64
                R[...] = R - green_gauss(part, Fx.astype(F32) if round_flux else Fx, dim)
        R = np.zeros_like(Q0)
        c.odom(f, Q0.copy(), R)
        return R

    Ro, R32 = residual(False), residual(True)
    scale = np.abs(Ro).max(axis=0)
    rel = (np.abs(R32 - Ro) / np.maximum(np.abs(Ro), 1e-3 * scale)).max(axis=1)
    assert (rel > 1e-5).mean() > 0.9, (rel > 1e-5).mean()
    assert (np.abs(R32 - Ro) / scale).max() < 1e-3          # ... while looking perfectly fine at the scale of the field


def test_one_fma_contraction_already_breaks_the_tolerance(get_case, ib, oracle):
    """The other half of DESIGN.md 4.1: Julia does not contract `a * b + c`; a GPU compiler does by default.  Fusing a
    SINGLE multiply-add of the reference -- the sensor blend of MUSCL, `uL * Df + (1 - Df) * uf`
    (src/ImmersedBoundary.jl:1150-1153), evaluated here with one rounding instead of two -- already moves the residual
    of a large share of the cells by more than 1e-5 relative.  Hence `-fmad=false` and the scalar adds behind the
    packed products."""
    from oracle import cfd
    import oracle.domain as od
    c = get_case("sphere3d_stl", 20_000)
    fl = cfd.Fluid()
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c.odom.centers))

    def muscl_fused_blend(part, u, du, dim, D):
        down, dnei = od.owner_distance(part, dim), od.neighbor_distance(part, dim)
        uo, un = od.at_owners(part, u, dim), od.at_neighbors(part, u, dim)
        col = lambda a: od._col(a, uo)
        gf = (un - uo) / col(down + dnei)
        duo, dun = od.at_owners(part, du, dim), od.at_neighbors(part, du, dim)
        s = od.minmod((2 * dun - gf) * col(dnei), (2 * duo - gf) * col(down))
        Df = col(np.maximum(np.maximum(od.at_owners(part, D, dim), od.at_neighbors(part, D, dim)), F32(1e-7)))
        t = (F32(1.0) - Df) * ((uo * col(dnei) + un * col(down)) / col(down + dnei))
        fma = lambda a, b, c_: (a.astype(np.float64) * b.astype(np.float64) + c_.astype(np.float64)).astype(F32)   # one rounding
        return fma(uo + s, Df + 0 * uo, t), fma(un - s, Df + 0 * uo, t)

    def residual(fused):
        def f(part, Q, R):
            P = cfd.state2primitive(fl, Q)
            D = od.JST_sensor(part, P[:, 0])
            R[...] = 0
            for dim in range(part.ndims):
                g = od.cell_gradient(part, P, dim)
                PL, PR = muscl_fused_blend(part, P, g, dim, D) if fused else od.MUSCL(part, P, g, dim, D=D, high_order=False)
                R[...] = R - od.green_gauss(part, cfd.inviscid_fluxes_hll(fl, PL, PR, dim), dim)
        R = np.zeros_like(Q0)
        c.odom(f, Q0.copy(), R)
        return R

    Ro, Rf = residual(False), residual(True)
    scale = np.abs(Ro).max(axis=0)
    rel = (np.abs(Rf - Ro) / np.maximum(np.abs(Ro), 1e-3 * scale)).max(axis=1)
    assert (rel > 1e-5).mean() > 0.2, (rel > 1e-5).mean()
