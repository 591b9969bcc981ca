"""Committed fixtures (tests/golden/oracle_fixtures.npz, produced by make_fixtures.py from the oracle):
CPU: the oracle and the product's host builder still reproduce them; GPU: the product reproduces them without the oracle."""
import os

import numpy as np
import pytest

F32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fixtures.npz"))


def _adv_mesh(ib):
    seg = lambda a, b: ib.Stereolitography(np.array([a, b], dtype=np.float64))
    h = F32(1e-2)
    return ib.Mesh([0.0, 0.0], [1.0, 1.0], ("lower", seg([0., 0.], [1., 0.]), h), ("upper", seg([0., 0.], [0., 1.]), h),
                   refinement_regions=[(ib.Line([0.0, 0.0], [1.0, 1.0]), F32(2) * h), (ib.Line([0.0, 0.0], [0.5, 0.5]), h)])


def test_builder_reproduces_fixtures(ib, get_case):
    msh = _adv_mesh(ib)
    assert np.array_equal(msh.block_origins, FIX["adv_block_origins"]) and np.array_equal(msh.block_widths, FIX["adv_block_widths"])
    dom = ib.Domain(msh, hypercube_families=[("outlet", [(0, True), (1, True)])], upload=False)
    assert np.array_equal(dom.faces(), FIX["adv_faces"])
    for name in ("outlet", "lower", "upper"):
        assert np.array_equal(dom.boundaries[name][1].ghost_indices, FIX[f"adv_ghosts_{name}"])
    c = get_case("advection")                       # the oracle has not drifted either
    assert np.array_equal(c.odom.faces, FIX["adv_faces"])
    m3 = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12)))
    assert np.array_equal(m3.block_origins, FIX["sph_block_origins"]) and np.array_equal(m3.block_widths, FIX["sph_block_widths"])


@pytest.mark.gpu
def test_gpu_reproduces_fixtures(ib):
    msh = _adv_mesh(ib)
    dom = ib.Domain(msh, hypercube_families=[("outlet", [(0, True), (1, True)])], build_partitions=False)
    N = len(dom)
    ud, sp = ib.DeviceArray(N, 1, True), ib.DeviceArray(N, 1, True)
    rng = np.random.default_rng(7)
    u = rng.random(N).astype(F32)
    C = (0.5 + rng.random((N, 2))).astype(F32)
    ib.residual_advection(dom, ib.DeviceArray.from_host(u), ib.DeviceArray.from_host(C), ud, sp)
    assert np.abs(ud.to_host() - FIX["adv_ud"]).max() <= 2e-5 * np.abs(FIX["adv_ud"]).max()
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    m3 = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12)))
    d3 = ib.Domain(m3, hypercube_families=fams, build_partitions=False)
    assert np.array_equal(d3.boundaries["wall"][1].ghost_indices, FIX["sph_ghosts_wall"])
    R, cf = ib.DeviceArray(len(d3), 5, False), ib.DeviceArray(len(d3), 1, True)
    Q = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(d3.cells()[0]))
    ib.residual_euler(d3, ib.Fluid(), ib.DeviceArray.from_host(Q), R, cf)
    Rg, cg = R.to_host(), cf.to_host()
    assert np.array_equal(Rg[::16], FIX["sph_R_sub"]) and np.array_equal(cg[::16], FIX["sph_cfl_sub"])  # bit for bit
    # checksums over ALL cells (summation order differs between the layouts, hence a tolerance far below 1 ulp of float32)
    assert np.allclose(np.abs(Rg.astype(np.float64)).sum(axis=0), FIX["sph_R_abs"], rtol=1e-9, atol=0)
    assert np.allclose(Rg.astype(np.float64).sum(axis=0), FIX["sph_R_sum"], rtol=0, atol=1e-9 * FIX["sph_R_abs"].max())
    assert np.isclose(cg.astype(np.float64).sum(), FIX["sph_cfl_sum"], rtol=1e-10)


def test_pinning_kit_file_format_round_trip(tmp_path):
    """tools/gen_reference_fixtures.jl (for a Julia owner) and tests/test_reference_fixtures.py share a file format; with
    no Julia here, exercise the consumer on files written FROM THE ORACLE in that format (a format test, not parity)."""
    import subprocess
    import sys
    env = dict(os.environ, IBX_REFERENCE_FIXTURES=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fake_reference_fixtures.py"), str(tmp_path)],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_reference_fixtures.py"), "-q", "-x"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0 and "3 passed" in r.stdout, r.stdout[-2000:]
