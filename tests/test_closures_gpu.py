"""Pointwise closure kernels (csrc/closures.cu) through the C ABI against the oracle on the same seeded inputs.

Tolerance: the kernels follow the reference's Float32 operation order, so most results are bit-identical; where the
reference leaves Float32 (pow through Float64, log / exp) a last-bit difference of the transcendental is allowed:
1e-6 relative (1e-5 for the 20-step wall-function fixed point)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32
N = 20_000


def _dev(ib, a):
    return ib.DeviceArray.from_host(np.asfortranarray(a))


def _close(a, b, rtol=1e-6):
    a, b = np.asarray(a), np.asarray(b)
    return np.allclose(a, b.reshape(a.shape), rtol=rtol, atol=1e-30)


def _state(rng, nd):
    P = np.empty((N, 2 + nd), F32)
    P[:, 0] = rng.uniform(5e4, 2e5, N)
    P[:, 1] = rng.uniform(5.0, 1500.0, N)          # includes values below the 10 K clamp
    P[:, 2:] = rng.uniform(-300, 300, (N, nd))
    return P


@pytest.mark.parametrize("nd", [2, 3])
def test_transport_and_viscous_fluxes(ib, oracle, nd):
    rng = np.random.default_rng(100 + nd)
    cfd = oracle.cfd
    fl, ofl = ib.Fluid(), cfd.Fluid()
    P = _state(rng, nd)
    Pg = [rng.standard_normal((N, 2 + nd)).astype(F32) * F32(50) for _ in range(nd)]
    T = _dev(ib, P[:, 1].copy())
    assert _close(ib.dynamic_viscosity(fl, T).to_host(), cfd.dynamic_viscosity(ofl, P[:, 1]))
    assert np.array_equal(ib.heat_conductivity(fl, T).to_host().ravel(), cfd.heat_conductivity(ofl, P[:, 1]))
    dP, dPg = _dev(ib, P), [_dev(ib, g) for g in Pg]
    for d in range(nd):
        assert _close(ib.viscous_fluxes(fl, dP, dPg, d).to_host(), cfd.viscous_fluxes(ofl, P, Pg, d), 2e-6)
    nrm = rng.standard_normal((N, nd)).astype(F32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(F32)
    mut = rng.uniform(0, 1e-3, N).astype(F32)
    got = ib.viscous_fluxes(fl, dP, dPg, _dev(ib, nrm), mu_t=_dev(ib, mut)).to_host()
    ref = cfd.viscous_fluxes(ofl, P, Pg, nrm, mu_t=mut)
    scale = np.abs(ref).max(axis=0) + 1e-30
    assert (np.abs(got - ref) / scale).max() < 2e-6
    assert _close(ib.viscous_fluxes(fl, dP, dPg, 0, mu_t=2.5e-4).to_host(), cfd.viscous_fluxes(ofl, P, Pg, 0, mu_t=F32(2.5e-4)), 2e-6)
    Cp = ib.pressure_coefficient(fl, _dev(ib, P[:, 0].copy()), 101325.0, 0.73).to_host().ravel()
    assert np.array_equal(Cp, cfd.pressure_coefficient(ofl, P[:, 0], 101325.0, 0.73))
    Pinf = [101325.0, 288.15, 230.0, 10.0][:2 + nd] if nd == 2 else [101325.0, 288.15, 230.0, 10.0, -4.0]
    assert abs(ib.Reynolds_number(fl, Pinf, 2.0) / cfd.reynolds_number(ofl, Pinf, 2.0) - 1) < 1e-6
    assert abs(ib.adjust_Reynolds(fl, Pinf, 2.0, 6.5e6).mu_ref / cfd.adjust_reynolds(ofl, Pinf, 2.0, 6.5e6).mu_ref - 1) < 1e-6
    with pytest.raises(ib.IbxError):
        ib.viscous_fluxes(fl, dP, dPg, nd)          # axis out of range


@pytest.mark.parametrize("nd", [2, 3])
def test_sensors_and_turbulence_closures(ib, oracle, nd):
    rng = np.random.default_rng(200 + nd)
    cfd, tb = oracle.cfd, oracle.turbulence
    T = ib.turbulence
    g = [[(rng.standard_normal(N) * 10).astype(F32) for _ in range(nd)] for _ in range(nd)]
    dg = [[_dev(ib, g[i][j]) for j in range(nd)] for i in range(nd)]
    assert np.array_equal(ib.shock_sensor(dg).to_host().ravel(), cfd.shock_sensor(g))
    assert np.array_equal(T.Ducros_sensor(dg).to_host().ravel(), tb.ducros_sensor(g))
    assert np.array_equal(T.shear_rate(dg).to_host().ravel(), tb.shear_rate(g))
    a, b, c = [(rng.uniform(9e4, 1.1e5, N)).astype(F32) for _ in range(3)]
    assert np.array_equal(ib.JST_sensor_3pt(_dev(ib, a), _dev(ib, b), _dev(ib, c)).to_host().ravel(), cfd.jst_sensor_3pt(a, b, c))
    Delta = rng.uniform(1e-3, 1e-1, N).astype(F32)
    S = tb.shear_rate(g)
    assert np.array_equal(T.Smagorinsky_nuSGS(_dev(ib, Delta), _dev(ib, S)).to_host().ravel(), tb.smagorinsky(Delta, S))
    if nd == 3:
        assert _close(T.WALE_nuSGS(_dev(ib, Delta), dg).to_host(), tb.wale(Delta, g), 2e-6)
    else:
        with pytest.raises(ValueError):
            T.WALE_nuSGS(_dev(ib, Delta), dg)
    Rey = np.concatenate([10 ** rng.uniform(-6, 7, N - 3), [0.0, -5.0, 1.0]]).astype(F32)
    got, ref = T.wall_function(_dev(ib, Rey)), tb.wall_function_rey(Rey)
    for k in ref:
        assert _close(got[k].to_host(), ref[k], 1e-5), k
    y, u, nu = rng.uniform(1e-5, 1e-2, N).astype(F32), rng.uniform(0.1, 100, N).astype(F32), rng.uniform(1e-5, 2e-5, N).astype(F32)
    got, ref = T.wall_function(_dev(ib, y), _dev(ib, u), _dev(ib, nu), n_iter=25, kappa=0.4), tb.wall_function(y, u, nu, n_iter=25, kappa=F32(0.4))
    for k in ref:
        assert _close(got[k].to_host(), ref[k], 2e-5), k
    kk, ee = rng.uniform(1e-3, 10, N).astype(F32), rng.uniform(1e-3, 10, N).astype(F32)
    got, ref = T.standard_keps(_dev(ib, kk), _dev(ib, ee), _dev(ib, S)), tb.standard_keps(kk, ee, S)
    for k in ref:
        assert np.array_equal(got[k].to_host().ravel(), ref[k]), k
    R = rng.uniform(1e-6, 1e-3, N).astype(F32)
    gR, gS = rng.standard_normal((N, nd)).astype(F32), rng.standard_normal((N, nd)).astype(F32)
    got, ref = T.Wray_Agarwal(_dev(ib, R), _dev(ib, S), _dev(ib, gR), _dev(ib, gS)), tb.wray_agarwal(R, S, gR, gS)
    assert np.array_equal(got["nu_R"].to_host().ravel(), ref["nu_R"]) and np.array_equal(got["S"].to_host().ravel(), ref["S"])
