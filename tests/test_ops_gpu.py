"""Per-operator CUDA kernels (through the C ABI) against the oracle on the same seeded inputs.

ops.cu/cfd.cu are compiled without FMA contraction and mirror the reference's operation order, so every
operator is expected to be BIT-EXACT against the float32 oracle; the only tolerance is on HLL (the reference
promotes to Float64 there and the device stores float32) and on reductions (summation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def _rand(n, cols=None, seed=0, lo=0.5, hi=1.5):
    rng = np.random.default_rng(seed)
    shape = (n,) if cols is None else (n, cols)
    return (lo + (hi - lo) * rng.random(shape)).astype(F32)


@pytest.mark.parametrize("name,mps", [("advection", 100_000), ("advection", 3_000), ("rae2822", 10_000), ("sphere3d", 40_000)])
def test_grid_operators_bit_exact(get_case, ib, oracle, name, mps):
    c = get_case(name, mps, upload=True)
    D = oracle.domain
    for pid in sorted(c.odom.partitions)[:3]:
        op, part = c.odom.partitions[pid], c.dom.partitions[pid]
        n = len(op.domain)
        u1, u3, du3 = _rand(n, seed=pid), _rand(n, 3, seed=pid + 1), _rand(n, 3, seed=pid + 2, lo=-1, hi=1)
        Dv = _rand(n, seed=pid + 3, lo=0.0, hi=1.0)
        d1, d3, dd3, dD = (ib.DeviceArray.from_host(a) for a in (u1, u3, du3, Dv))
        assert np.array_equal(part.spacing.to_host(), op.spacing) and np.array_equal(part.centers.to_host(), op.centers)
        for dim in range(c.dom.ndims):
            for u, du_ in ((u1, d1), (u3, d3)):
                assert np.array_equal(ib.at_owners(part, du_, dim).to_host(), D.at_owners(op, u, dim))
                assert np.array_equal(ib.at_neighbors(part, du_, dim).to_host(), D.at_neighbors(op, u, dim))
                uf = D.at_faces(op, u, dim)
                duf = ib.at_faces(part, du_, dim)
                assert np.array_equal(duf.to_host(), uf)
                assert np.array_equal(ib.green_gauss(part, duf, dim).to_host(), D.green_gauss(op, uf, dim))
                assert np.array_equal(ib.unsigned_green_gauss(part, duf, dim).to_host(), D.unsigned_green_gauss(op, uf, dim))
                assert np.array_equal(ib.cell_gradient(part, du_, dim).to_host(), D.cell_gradient(op, u, dim))
                assert np.array_equal(ib.face_gradient(part, du_, dim).to_host(), D.face_gradient(op, u, dim))
                assert np.array_equal(ib.JST_sensor(part, du_, dim).to_host(), D.JST_sensor(op, u, dim))
            assert np.array_equal(ib.face_distance(part, dim).to_host(), D.face_distance(op, dim))
            assert np.array_equal(ib.owner_distance(part, dim).to_host(), D.owner_distance(op, dim))
            assert np.array_equal(ib.neighbor_distance(part, dim).to_host(), D.neighbor_distance(op, dim))
            for kw, dkw in (({}, {}), ({"D": Dv}, {"D": dD}), ({"D": Dv, "high_order": True}, {"D": dD, "high_order": True})):
                L, R = D.MUSCL(op, u3, du3, dim, **kw)
                dL, dR = ib.MUSCL(part, d3, dd3, dim, **dkw)
                assert np.array_equal(dL.to_host(), L) and np.array_equal(dR.to_host(), R)
        assert np.array_equal(ib.JST_sensor(part, d1).to_host(), D.JST_sensor(op, u1))
        g = ib.cell_gradient(part, d1)
        og = D.cell_gradient(op, u1)
        assert all(np.array_equal(a.to_host(), b) for a, b in zip(g, og))
        fg = ib.face_gradient(part, d1, 0, g)
        ofg = D.face_gradient(op, u1, 0, og)
        assert all(np.array_equal(a.to_host(), b) for a, b in zip(fg, ofg))
        dv = ib.divergent(part, tuple(ib.at_faces(part, d3, dim) for dim in range(c.dom.ndims)))
        odv = D.divergent(op, tuple(D.at_faces(op, u3, dim) for dim in range(c.dom.ndims)))
        assert np.array_equal(dv.to_host(), odv)


def test_cfd_kernels(ib, oracle):
    cfd = oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    rng = np.random.default_rng(5)
    n = 5000
    for nd in (2, 3):
        P = np.concatenate([_rand(n, 1, 1, 5e4, 2e5), _rand(n, 1, 2, 5.0, 400.0), _rand(n, nd, 3, -400, 400)], axis=1)
        PR = (P * (1 + 0.05 * (rng.random(P.shape) - 0.5))).astype(F32)
        dP, dPR = ib.DeviceArray.from_host(P), ib.DeviceArray.from_host(PR)
        Q = cfd.primitive2state(ofl, P)
        dQ = ib.primitive2state(fl, dP)
        assert np.array_equal(dQ.to_host(), Q)
        assert np.array_equal(ib.state2primitive(fl, dQ).to_host(), cfd.state2primitive(ofl, Q))
        assert np.array_equal(ib.speed_of_sound(fl, dP.col(1)).to_host(), cfd.speed_of_sound(ofl, P[:, 1]))
        nu = _rand(n, seed=9, lo=0, hi=1)
        dnu = ib.DeviceArray.from_host(nu)
        for dim in range(nd):
            F = cfd.inviscid_fluxes_hll(ofl, P, PR, dim)
            dF = ib.inviscid_fluxes(fl, dP, dPR, dim)
            assert dF.f64 and np.array_equal(dF.to_host(), F)  # Float64 like the reference (src/cfd.jl:504-507)
            dF32 = ib.DeviceArray(n, nd + 2, False)
            ib._lib.call("ibx_inviscid_fluxes_hll", ib.context(), fl.c, dP.h, dPR.h, dim, dF32.h)
            assert np.array_equal(dF32.to_host(), F.astype(F32))  # float32 output: same evaluation rounded once
            Fs = cfd.inviscid_fluxes_sensor(ofl, P, PR, nu, nu * F32(0.5), dim)
            dFs = ib.inviscid_fluxes(fl, dP, dPR, dnu, dnu * 0.5, dim).to_host()
            assert np.array_equal(dFs, Fs)
        nrm = rng.standard_normal((n, nd)).astype(F32)
        nrm /= np.linalg.norm(nrm, axis=1)[:, None].astype(F32)
        dn = ib.DeviceArray.from_host(nrm)
        for Pinf, nf in ((np.array([101325.0, 288.15] + [150.0, 20.0, -5.0][:nd], F32), False),
                         (np.array([101325.0, 288.15] + [700.0, 20.0, -5.0][:nd], F32), False),
                         (np.array([101325.0, 288.15, 0.0], F32), True)):
            ob = cfd.FlowBC(ofl, Pinf, normal_flow=nf)(P, nrm)
            db = ib.FlowBC(fl, Pinf, normal_flow=nf)(dP, dn).to_host()
            assert np.array_equal(db, ob)
            # keyword arguments of src/cfd.jl:245-249: transpiration and the wall-shear scaling
            dist, dudn = _rand(n, seed=11, lo=1e-3, hi=1e-2), _rand(n, seed=12, lo=0, hi=5e3)
            tr = _rand(n, seed=13, lo=-1, hi=1)
            ob = cfd.FlowBC(ofl, Pinf, normal_flow=nf)(P, nrm, image_distances=dist, du_dn=dudn, transpiration=tr)
            db = ib.FlowBC(fl, Pinf, normal_flow=nf)(dP, dn, image_distances=ib.DeviceArray.from_host(dist),
                                                     du_dn=ib.DeviceArray.from_host(dudn), transpiration=ib.DeviceArray.from_host(tr)).to_host()
            assert np.allclose(db, ob, rtol=2e-7, atol=0) and np.array_equal(db[:, :2], ob[:, :2])
            ob = cfd.FlowBC(ofl, Pinf, normal_flow=nf)(P, nrm, transpiration=F32(0.25))
            assert np.array_equal(ib.FlowBC(fl, Pinf, normal_flow=nf)(dP, dn, transpiration=0.25).to_host(), ob)
    with pytest.raises(ib.IbxError, match="passed together"):
        ib.FlowBC(fl, Pinf, normal_flow=True)(dP, dn, du_dn=dn.col(0))
    with pytest.raises(ib.IbxError, match="Only 3 parcels"):
        ib.FlowBC(fl, np.zeros(4, F32), normal_flow=True)(dP, dn)


def test_accumulator_and_elementwise(ib, oracle):
    rng = np.random.default_rng(11)
    n, m = 3000, 700
    lens = rng.integers(0, 9, size=m)
    inds = [rng.integers(0, n, size=l) for l in lens]
    ws = [rng.standard_normal(l).astype(F32) for l in lens]
    v = rng.standard_normal((n, 3)).astype(F32)
    acc, oacc = ib.Accumulator.from_lists(inds, ws), oracle.accumulator.Accumulator(inds, ws, first_index=True)
    assert np.array_equal(acc(v), oacc(v))                     # host arrays in -> host arrays out
    assert np.array_equal(acc(v[:, 0].copy()), oacc(v[:, 0].copy()))
    inds2 = [rng.integers(0, m, size=max(l, 1)) for l in lens]
    ws2 = [rng.standard_normal(max(l, 1)).astype(F32) for l in lens]
    v2 = rng.standard_normal(m).astype(F32)
    assert np.array_equal(ib.Accumulator.from_lists(inds2, ws2)(v2, delta=True),
                          oracle.accumulator.Accumulator(inds2, ws2)(v2, delta=True))
    un, oun = ib.Accumulator.from_lists(inds2), oracle.accumulator.Accumulator(inds2)
    assert np.array_equal(un(v2), oun(v2))
    # the `f` and `op` keyword arguments (src/accumulator.jl:78-111), weighted with and without Delta, and unweighted
    a2, oa2 = ib.Accumulator.from_lists(inds2, ws2), oracle.accumulator.Accumulator(inds2, ws2)
    fs = {"abs": np.abs, "abs2": lambda x: x * x, "sign": np.sign}
    ops = {"+": None, "max": np.maximum, "min": np.minimum, "*": lambda p, q: p * q}
    for fname, f in fs.items():
        for oname, op in ops.items():
            for delta in (False, True):
                assert np.array_equal(a2(v2, delta=delta, f=fname, op=oname), oa2(v2, delta=delta, f=f, op=op)), (fname, oname, delta)
            assert np.array_equal(un(v2, f=fname, op=oname), oun(v2, f=f, op=op)), (fname, oname)
    assert np.array_equal(acc(v, f="abs", op="max"), oacc(v, f=np.abs, op=np.maximum))     # empty rows stay 0
    kat = ib.Accumulator.from_lists([[0, 1], [1, 2, 3]], [[-1.0, 2.0], [3.0, 4.0, 5.0]])
    assert np.array_equal(kat(np.array([1, 2, 3, 4], F32)), np.array([3.0, 38.0], F32))  # src/accumulator.jl:25-34
    a, b = ib.DeviceArray.from_host(v), ib.DeviceArray.from_host(v[::-1].copy())
    s = ib.DeviceArray.from_host(v[:, 0].copy())
    assert np.array_equal(((a + b) * s / 2 + abs(s) * (a - b) / 2).to_host(),
                          (v + v[::-1]) * v[:, :1] / 2 + np.abs(v[:, :1]) * (v - v[::-1]) / 2)
    assert np.array_equal(ib.maximum(a, b).to_host(), np.maximum(v, v[::-1]))
    assert np.array_equal((1.0 / (a * a + 1.0)).to_host(), F32(1) / (v * v + F32(1)))
    assert np.isclose(a.sum(), v.astype(np.float64).sum()) and a.max() == v.max() and a.min() == v.min()
    vn = v.copy()
    vn[17, 1] = np.nan                                  # Julia's maximum / minimum propagate NaN: so do the reductions
    an = ib.DeviceArray.from_host(vn)
    assert np.isnan(an.max()) and np.isnan(an.min()) and np.isnan(an.sum())
    assert np.isclose(a.norm(), np.linalg.norm(v.astype(np.float64))) and a.maxabs() == np.abs(v).max()
    assert np.isclose(ib.dot(a, b), (v.astype(np.float64) * v[::-1]).sum())
    assert np.allclose(a.sum(per_column=True), v.astype(np.float64).sum(axis=0))


def test_dom_closure_advection_march(get_case, ib, oracle):
    """test/advection.jl run through dom(f, ...) with the closure written exactly like the reference's, on the
    device, against the same steps in the oracle: bit-exact after several steps, on 1 and on 4 partitions."""
    E, OD = oracle.euler, oracle.domain
    for mps in (100_000, 3_000):
        c = get_case("advection", mps, upload=True)
        N = len(c.dom)
        C = np.ones((N, 2), F32)
        dC = ib.DeviceArray.from_host(C)

        def residual(part, u, ud, Cl):
            Dn = ib.JST_sensor(part, u)
            for dim in range(part.ndims):
                Cf = ib.at_faces(part, Cl.col(dim), dim)
                gu = ib.cell_gradient(part, u, dim)
                uL, uR = ib.MUSCL(part, u, gu, dim, D=Dn, high_order=True)
                ud -= ib.green_gauss(part, (uL + uR) * Cf / 2 + abs(Cf) * (uL - uR) / 2, dim)

        def spectral(part, Cl):
            s = None
            for dim in range(part.ndims):
                t = ib.unsigned_green_gauss(part, ib.at_faces(part, Cl.col(dim), dim), dim)
                s = t if s is None else ib.maximum(s, t)
            return F32(0.5) / s.max()

        u, ou = np.zeros(N, F32), np.zeros(N, F32)
        for _ in range(5):
            dt = min(c.dom(spectral, dC)) * F32(0.75)
            odt = min(c.odom(lambda p, Cl: F32(0.5) / E.advection_spectral(p, Cl).max(), C.copy())) * F32(0.75)
            assert dt == odt
            ud, oud = np.zeros(N, F32), np.zeros(N, F32)
            c.dom(residual, u, ud, dC)
            c.odom(lambda p, u_, ud_, Cl: E.advection_residual(p, u_, ud_, Cl), ou, oud, C.copy())
            # the very first non-trivial residual sees identical inputs: bit-exact; afterwards the ghost values differ
            # in the last bit (interpolation weights: float32 SVD in the oracle, double Jacobi in the product)
            assert np.abs(ud - oud).max() <= 2e-5 * max(1.0, np.abs(oud).max())
            u += ud * dt
            ou += oud * odt
            for dm, mod, uu in ((c.dom, ib, u), (c.odom, OD, ou)):
                mod.impose_bc(lambda b, x: F32(1.0), dm, "upper", uu)
                mod.impose_bc(lambda b, x: F32(0.0), dm, "lower", uu)
                mod.impose_bc(lambda b, x: x.copy(), dm, "outlet", uu)
            assert np.abs(u - ou).max() < 1e-5
        assert u.max() > 0.5


def test_dissipation_and_volume_integral(get_case, ib, oracle):
    c = get_case("dissipation", upload=True)
    E, OD = oracle.euler, oracle.domain
    N = len(c.dom)
    uv = _rand(N, 2, seed=4)
    uvd, ouvd = np.zeros((N, 2), F32), np.zeros((N, 2), F32)

    def f(part, uv, uvd):
        for dim in range(part.ndims):
            uvd += ib.green_gauss(part, ib.face_gradient(part, uv, dim), dim)

    c.dom(f, uv.copy(), uvd)
    c.odom(E.dissipation_residual, uv.copy(), ouvd)
    assert np.array_equal(uvd, ouvd)
    vi, ovi = ib.volume_integral(c.dom, uv), OD.volume_integral(c.odom, uv)
    assert np.allclose(vi, ovi, rtol=1e-5)
    assert abs(ib.volume_integral(c.dom, np.ones(N, F32)) - 1.0) < 1e-5


def test_surface_and_bc_vector_return(get_case, ib, oracle):
    c = get_case("rae2822", upload=True)
    N = len(c.dom)
    u = _rand(N, 2, seed=8)
    s, os_ = c.dom.surfaces["wall"], c.odom.surfaces["wall"]
    assert np.abs(s(u) - os_(u)).max() < 1e-5 and np.abs(s.at_offset(u) - os_.at_offset(u)).max() < 1e-5
    si = ib.surface_integral(s, s(u))
    assert np.allclose(si, oracle.domain.surface_integral(os_, os_(u)), rtol=1e-4)
    ny, ony = np.zeros(N, F32), np.zeros(N, F32)  # test/rae2822.jl:31-34
    ib.impose_bc(lambda b, x: b.normals.col(1), c.dom, "wall", ny)
    oracle.domain.impose_bc(lambda b, x: b.normals[:, 1], c.odom, "wall", ony)
    assert np.array_equal(ny, ony)


def test_transfer_operators_and_solvers(get_case, ib, oracle):
    c = get_case("advection", upload=True)
    cd, pro, coa = ib.multigrid(c.dom)
    ocd, opro, ocoa = oracle.domain.multigrid(c.odom)
    P = _rand(len(c.dom), 4, seed=3)
    Pc, oPc = coa[0](P), ocoa[0](P)
    assert np.abs(Pc - oPc).max() < 1e-6
    assert np.abs(pro[0](Pc) - opro[0](oPc)).max() < 1e-6
    # point-implicit: block pinv + apply, then the linear solve drivers on a block-diagonal problem
    rng = np.random.default_rng(1)
    n, nv = 4000, 5
    A = (rng.random((n, nv, nv)) * 0.1 + np.eye(nv)[None] * 2).astype(F32)
    A[::7] = 0                                                     # singular blocks: pinv gives 0
    Dm = ib.DeviceArray.from_host(A.transpose(0, 2, 1).reshape(n, nv * nv))  # column j + nv * i
    prec = ib.PIPreconditioner(Dm, nv)
    v = rng.standard_normal((n, nv)).astype(F32)
    ref = np.einsum("pji,pi->pj", np.linalg.pinv(A.astype(np.float64)), v)
    assert np.abs(prec(ib.DeviceArray.from_host(v)).to_host() - ref).max() < 2e-5
    xs = rng.random((n, nv)).astype(F32)
    A2 = (rng.random((n, nv, nv)) * 0.1 + np.eye(nv)[None] * 2).astype(F32)
    dA = ib.DeviceArray.from_host(A2.transpose(0, 2, 1).reshape(n, nv * nv))
    dxs = ib.DeviceArray.from_host(xs)

    def f(x):
        out, dx_ = ib.DeviceArray(n, nv, False), x - dxs
        ib._lib.call("ibx_block_apply", ib.context(), dA.h, nv, dx_.h, out.h)
        return out

    lin, b, pre = ib.linearize(f, ib.DeviceArray(n, nv, False).fill(0.0), n_hutchinson_samples=6, h=1e-3)
    dx, ratio = ib.solve(lin, b, pre, n_iter=40, rtol=1e-3)
    assert ratio < 1e-2 and np.abs(dx.to_host() - xs).max() < 5e-2
    Q = ib.DeviceArray(n, nv, False).fill(0.0)
    ratio = ib.FAS(lambda l, Q: (-f(Q), 0.4), Q, n_iter=200, rtol=1e-3)
    assert ratio < 2e-3 and np.abs(Q.to_host() - xs).max() < 1e-2
    mg = ib.Multigrid(c.dom.cells()[0], 2)
    omg = oracle.mgrid.Multigrid(c.odom.centers, 2)
    w = _rand(len(c.dom), seed=6)
    assert np.abs(mg.prolongators[1](mg.coarseners[1](w)) - omg.prolongators[1](omg.coarseners[1](w))).max() < 1e-6
