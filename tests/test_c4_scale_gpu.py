"""The benched workload (C4: 3-D sphere octree, BASELINE.json configs[3]) at 3 M and 10 M cells: GPU step (IB ghost update
of both boundaries + fused Euler residual, through the C ABI) against the compiled CPU restatement of the reference
path (oracle/cpu_ref.c, pinned bit for bit to the NumPy oracle by tests/test_oracle_cpu_ref.py).

Same recipe as bench.py (box (-16)^3..16^3, sphere r = 0.5, refinement ball 0.75, growth ratio 2, block size 8), two
octree levels coarser (level 8 -> 3.07 M cells) and one level coarser (level 9 -> 10.3 M cells) than the benched level
10 (50.2 M cells).  Bit-exact: `array_equal` on the ghost-updated state, the residual and the CFL array."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def _mesh(ib, level, surface):
    h = F32(32.0 / 2 ** level / 8 * 1.01)
    return ib.Mesh([-16, -16, -16], [32, 32, 32], ("wall", surface, h), refinement_regions=[(ib.Ball([0, 0, 0], 0.75), h)])


@pytest.mark.parametrize("level,cells", [(8, 3_072_000), (9, 10_268_672)])
def test_c4_recipe_gpu_equals_cpu_reference(ib, oracle, level, cells):
    from oracle import cfd, cpu_ref
    msh = _mesh(ib, level, ib.Sphere([0, 0, 0], 0.5))
    assert len(msh) == cells
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    dom = ib.Domain(msh, max_partition_size=400_000, hypercube_families=fams, build_partitions=True, build_surfaces=False,
                    upload=True)
    ref = cpu_ref.CpuRef.from_builder(dom)
    fl, ofl = ib.Fluid(), cfd.Fluid()
    a = np.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)
    wall = np.array([101325.0, 288.15, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fl, wall, normal_flow=True)), ("farfield", ib.FlowBC(fl, Pinf))]
    obcs = [("wall", cfd.FlowBC(ofl, wall, normal_flow=True)), ("farfield", cfd.FlowBC(ofl, Pinf))]
    N = len(dom)
    Q0 = np.asfortranarray(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(dom.cells()[0])))
    Q = ib.DeviceArray.from_host(Q0)
    R, cf = ib.DeviceArray(N, 5, False), ib.DeviceArray(N, 1, True)
    ib.ghost_update_euler(dom, fl, Q, bcs)
    ib.residual_euler(dom, fl, Q, R, cf)
    Qo = Q0.copy(order="F")
    Ro, co = np.zeros((N, 5), F32, order="F"), np.zeros(N, F32)
    ref.ghost_update(ofl, Qo, obcs)
    ref.residual(ofl, Qo, Ro, co)
    Qg, Rg, cg = Q.to_host(), R.to_host(), cf.to_host()
    n_ghost = int((Qo != Q0).any(axis=1).sum())
    assert n_ghost > 10_000
    assert np.array_equal(Qg, Qo), f"ghost update differs in {(Qg != Qo).any(axis=1).sum()} of {n_ghost} ghost cells"
    assert np.array_equal(Rg, Ro), f"residual differs in {(Rg != Ro).any(axis=1).sum()} of {N} cells, max {np.abs(Rg - Ro).max()}"
    assert np.array_equal(cg, co)
    # the same evaluation through the host-buffer entry point the end-to-end number is measured on
    Rh, ch = ib.pinned_empty((N, 5)), ib.pinned_empty((N,))
    Qh = ib.pinned_empty((N, 5))
    Qh[...] = Q0
    ib.euler_step_host(dom, fl, bcs, Qh, Rh, ch)
    assert np.array_equal(Rh, Ro) and np.array_equal(ch, co)


def test_host_buffer_slots_first_use_on_a_large_mesh(ib):
    """First-ever use of BOTH host-buffer slots back to back on a multi-million-cell mesh (the slot arrays were
    cleared on the compute stream while the upload ran on the copy stream -- the clear could overtake the upload)."""
    msh = _mesh(ib, 8, ib.Sphere([0, 0, 0], 0.5))
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    dom = ib.Domain(msh, hypercube_families=fams, build_partitions=False, build_surfaces=False, upload=True)
    fl = ib.Fluid()
    a = np.sqrt(1.4 * 283.0 * 288.15)
    bcs = [("wall", ib.FlowBC(fl, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)),
           ("farfield", ib.FlowBC(fl, np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)))]
    N = len(dom)
    Q0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(dom.cells()[0]))
    Qp = [ib.pinned_empty((N, 5)) for _ in range(2)]
    Qp[0][...] = Q0
    Qp[1][...] = Q0
    Rs, cs = [ib.pinned_empty((N, 5)) for _ in range(2)], [ib.pinned_empty((N,)) for _ in range(2)]
    # a different mesh size first, so that both slots are re-allocated by the calls below
    small = ib.Domain(ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12))), hypercube_families=fams,
                      build_partitions=False, build_surfaces=False, upload=True)
    qs = np.asfortranarray(ib.synthetic.primitive2state_host(ib.synthetic.euler_state(small.cells()[0])))
    rs, c_s = np.zeros_like(qs), np.zeros(len(small), F32)
    for k in range(2):
        ib.euler_step_host_begin(small, fl, bcs, qs, rs, c_s, k)
        ib.euler_step_host_end(k)
    for rep in range(3):
        ib.euler_step_host_begin(dom, fl, bcs, Qp[0], Rs[0], cs[0], 0)
        ib.euler_step_host_begin(dom, fl, bcs, Qp[1], Rs[1], cs[1], 1)
        ib.euler_step_host_end(0)
        ib.euler_step_host_end(1)
        assert np.isfinite(Rs[0]).all() and np.isfinite(Rs[1]).all()
        assert np.array_equal(Rs[0], Rs[1]) and np.array_equal(cs[0], cs[1])
    Q = ib.DeviceArray.from_host(Q0)
    R, cf = ib.DeviceArray(N, 5, False), ib.DeviceArray(N, 1, True)
    ib.ghost_update_euler(dom, fl, Q, bcs)
    ib.residual_euler(dom, fl, Q, R, cf)
    assert np.array_equal(R.to_host(), Rs[0]) and np.array_equal(cf.to_host(), cs[0])
