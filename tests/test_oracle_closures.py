"""Oracle restatement of the pointwise closures (src/cfd.jl transport / viscous / sensors, src/turbulence.jl) against
analytic invariants -- the reference ships no test vector for them."""
import numpy as np

F32 = np.float32


def _vg(nd, n, fill):
    return [[np.full(n, fill(i, j), F32) for j in range(nd)] for i in range(nd)]


def test_transport_properties(oracle):
    cfd = oracle.cfd
    fl = cfd.Fluid()
    assert np.allclose(cfd.dynamic_viscosity(fl, np.array([fl.T_ref], F32)), fl.mu_ref, rtol=1e-6)
    T = np.array([5.0, 10.0, 300.0, 1500.0], F32)
    mu = cfd.dynamic_viscosity(fl, T)
    assert mu[0] == mu[1] and np.all(mu > 0)        # clamp at 10 K (the reference's exponent is 2/3, not Sutherland's 3/2)
    exact = fl.mu_ref * (300.0 / fl.T_ref) ** (2 / 3) * (fl.T_ref + fl.S) / (300.0 + fl.S)
    assert abs(mu[2] / exact - 1) < 1e-6
    k = cfd.heat_conductivity(fl, T)
    assert np.allclose(k, fl.k[0] + fl.k[1] * T, rtol=1e-6)
    assert cfd.pow32(np.array([2.0], F32), F32(0.5))[0] == F32(np.sqrt(2.0)) and cfd.pow32(np.array([3.0], F32), 3)[0] == F32(27)
    assert cfd.pressure_coefficient(fl, np.array([101325.0], F32), 101325.0, 0.5)[0] == 0
    Re = cfd.reynolds_number(fl, [101325.0, 288.15, 100.0, 0.0], 1.0)
    fl2 = cfd.adjust_reynolds(fl, [101325.0, 288.15, 100.0, 0.0], 1.0, 1e6)
    assert abs(cfd.reynolds_number(fl2, [101325.0, 288.15, 100.0, 0.0], 1.0) / 1e6 - 1) < 1e-5 and Re > 1e6


def test_viscous_fluxes_pure_shear_and_normal_form(oracle):
    cfd = oracle.cfd
    fl = cfd.Fluid()
    n, a, dTdy = 7, F32(3.0), F32(2.0)
    P = np.tile(np.array([101325.0, 300.0, 20.0, 0.0, 0.0], F32), (n, 1))
    Pg = [np.zeros((n, 5), F32) for _ in range(3)]
    Pg[1][:, 2] = a          # du/dy
    Pg[1][:, 1] = dTdy       # dT/dy
    F = cfd.viscous_fluxes(fl, P, Pg, 1)
    mu, k = cfd.dynamic_viscosity(fl, P[:, 1]), cfd.heat_conductivity(fl, P[:, 1])
    assert np.allclose(F[:, 2], mu * a) and np.allclose(F[:, 3], 0) and np.allclose(F[:, 0], 0)
    assert np.allclose(F[:, 1], k * dTdy + mu * a * P[:, 2], rtol=1e-6)
    rng = np.random.default_rng(0)
    Pg = [rng.standard_normal((n, 5)).astype(F32) for _ in range(3)]
    for d in range(3):
        e = np.zeros((n, 3), F32)
        e[:, d] = 1
        assert np.allclose(cfd.viscous_fluxes(fl, P, Pg, e, mu_t=F32(1e-5)), cfd.viscous_fluxes(fl, P, Pg, d, mu_t=F32(1e-5)), rtol=1e-5, atol=1e-9)
    F = cfd.viscous_fluxes(fl, P, Pg, 0)
    divu = Pg[0][:, 2] + Pg[1][:, 3] + Pg[2][:, 4]
    assert np.allclose(F[:, 2], mu * (2 * Pg[0][:, 2] - 2 / 3 * divu), rtol=1e-4, atol=1e-9)   # tau_xx


def test_sensors(oracle):
    cfd, tb = oracle.cfd, oracle.turbulence
    n = 5
    dil = _vg(3, n, lambda i, j: 2.0 if i == j else 0.0)
    rot = _vg(3, n, lambda i, j: {(1, 0): 3.0, (0, 1): -3.0}.get((i, j), 0.0))
    assert np.allclose(cfd.shock_sensor(dil), 1) and np.allclose(tb.ducros_sensor(dil), 1)
    assert cfd.shock_sensor(rot).max() < 1e-12 and tb.ducros_sensor(rot).max() < 1e-7
    rot2 = _vg(2, n, lambda i, j: {(1, 0): 3.0, (0, 1): -3.0}.get((i, j), 0.0))
    mix = [[rot2[i][j] + (F32(1.0) if i == j else F32(0)) for j in range(2)] for i in range(2)]
    assert np.allclose(cfd.shock_sensor(mix), 4.0 / (4.0 + 2 * 36.0))        # 2-D: the vorticity is counted twice (reference quirk)
    assert np.allclose(tb.ducros_sensor(mix), 4.0 / (4.0 + 36.0), rtol=1e-6)
    lin = np.array([1.0, 2.0, 3.0], F32)
    assert cfd.jst_sensor_3pt(lin - 1, lin, lin + 1).max() < 1e-13            # linear data: no second difference
    assert np.allclose(cfd.jst_sensor_3pt(lin, lin + 1, lin), 1)
    shear = _vg(3, n, lambda i, j: 4.0 if (i, j) == (0, 1) else 0.0)
    assert np.allclose(tb.shear_rate(shear), 4.0)
    D = np.full(n, 0.1, F32)
    assert np.allclose(tb.smagorinsky(D, tb.shear_rate(shear)), (0.17 * 0.1) ** 2 * 4.0, rtol=1e-6)
    assert tb.wale(D, shear).max() == 0                                         # WALE vanishes in pure shear
    rng = np.random.default_rng(1)
    g = [[rng.standard_normal(n).astype(F32) for _ in range(3)] for _ in range(3)]
    w = tb.wale(D, g)
    assert np.all(w > 0) and np.all(np.isfinite(w))


def test_wall_function_and_transport_models(oracle):
    tb = oracle.turbulence
    Rey = np.array([1e-3, 1.0, 1e2, 1e4, 1e6], F32)
    nt = tb.wall_function_rey(Rey)
    assert np.allclose(nt["y_plus"] * nt["u_plus"], Rey, rtol=1e-5)
    assert np.allclose(nt["y_plus"][:2], np.sqrt(Rey[:2]), rtol=1e-4)                      # viscous sublayer: u+ = y+
    yp = nt["y_plus"][-1]
    assert abs(nt["u_plus"][-1] - (np.log(yp) / 0.41 + 4.9)) < 1e-2                          # log layer
    assert np.all(nt["dudy_plus"] <= 1) and np.all(nt["mu_plus"] >= 0)
    y, u, nu = np.full(3, 1e-3, F32), np.array([1.0, 10.0, 50.0], F32), np.full(3, 1.5e-5, F32)
    wf = tb.wall_function(y, u, nu)
    assert np.allclose(wf["eps"], F32(0.09) * wf["omega"] * wf["k"]) and np.all(wf["u_tau"] > 0)
    k, eps, S = np.array([1.0, 2.0], F32), np.array([0.5, 4.0], F32), np.array([3.0, 1.0], F32)
    ke = tb.standard_keps(k, eps, S)
    assert np.allclose(ke["nu_t"], 0.09 * k ** 2 / eps) and np.allclose(ke["Sk"], ke["nu_t"] * S ** 2 - eps, rtol=1e-6)
    R = np.array([1e-4, 2e-4], F32)
    wa = tb.wray_agarwal(R, S, np.ones((2, 3), F32), np.ones((2, 3), F32) * F32(1e6))
    assert np.allclose(wa["S"], 10 * R) and np.allclose(wa["nu_R"], R * F32(0.72))           # source clipped at 10 R
