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
