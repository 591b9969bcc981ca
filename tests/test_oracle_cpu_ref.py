"""The compiled CPU baseline (oracle/cpu_ref.c: one OpenMP task per partition, array-at-a-time operators) is the SAME
restatement as the NumPy oracle: bit-identical residual, CFL term and ghost update on 2-D and 3-D meshes, and
independent of the thread count.  (Test infrastructure checking test infrastructure: bench.py times this library as
the CPU arm.)"""
import numpy as np
import pytest

F32 = np.float32


def _bcs(cfd, fl, nd):
    a = np.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0][:2 + nd], F32)
    return [("wall", cfd.FlowBC(fl, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)), ("farfield", cfd.FlowBC(fl, Pinf))]


@pytest.mark.parametrize("name,mps", [("rae2822", 10_000), ("sphere3d_stl", 20_000)])
def test_compiled_cpu_baseline_equals_numpy_oracle(get_case, ib, oracle, name, mps):
    from oracle import cpu_ref
    c = get_case(name, mps)
    E, cfd = oracle.euler, oracle.cfd
    fl = cfd.Fluid()
    odom = c.odom
    N, nd = odom.centers.shape
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(odom.centers))
    ref = cpu_ref.CpuRef.from_oracle(odom)
    # ghost update
    Qo = Q0.copy()
    g = E.euler_ghost_update(odom, fl, Qo, _bcs(cfd, fl, nd))
    Qc = np.asfortranarray(Q0)
    ref.ghost_update(fl, Qc, _bcs(cfd, fl, nd), n_threads=3)
    assert len(g) > 0 and np.array_equal(Qc, Qo)
    # residual on the ghost-updated state
    Ro, co = np.zeros_like(Qo), np.zeros(N, F32)
    odom(E.euler_residual(fl), Qo.copy(), Ro, co)
    out = []
    for nt in (1, 4):
        Rc, cc = np.zeros((N, nd + 2), F32, order="F"), np.zeros(N, F32)
        ref.residual(fl, Qc, Rc, cc, n_threads=nt)
        out.append((Rc, cc))
        assert np.array_equal(Rc, Ro), (nt, int((Rc != Ro).sum()))
        assert np.array_equal(cc, co)
    assert np.array_equal(out[0][0], out[1][0])


def test_cpu_baseline_on_builder_tables_equals_oracle_tables(get_case, ib, oracle):
    """bench.py feeds the compiled CPU baseline with the partition / boundary tables of the host-side C++ builder (the
    NumPy builder is too slow at bench size).  Same tables -> same bits as with the oracle's own tables; the donor
    weights are the one place the two builders differ (float32 SVD vs double Jacobi, <= 2e-6), visible only in the
    ghost update."""
    from oracle import cpu_ref
    c = get_case("rae2822", 10_000)
    cfd = oracle.cfd
    fl = cfd.Fluid()
    N, nd = c.odom.centers.shape
    Q0 = np.asfortranarray(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(c.odom.centers)))
    a, b = cpu_ref.CpuRef.from_oracle(c.odom), cpu_ref.CpuRef.from_builder(c.dom)
    out = []
    for ref in (a, b):
        R, cf = np.zeros((N, nd + 2), F32, order="F"), np.zeros(N, F32)
        ref.residual(fl, Q0, R, cf, n_threads=2)
        out.append((R, cf))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    Qa, Qb = Q0.copy(order="F"), Q0.copy(order="F")
    a.ghost_update(fl, Qa, _bcs(cfd, fl, nd))
    b.ghost_update(fl, Qb, _bcs(cfd, fl, nd))
    touched = (Qa != Q0).any(axis=1)
    assert touched.sum() > 0 and np.array_equal(touched, (Qb != Q0).any(axis=1))
    assert (np.abs(Qa - Qb) / np.abs(Qa).max(axis=0)).max() < 2e-6
