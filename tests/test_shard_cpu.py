"""Multi-rank host logic on CPU: world_size-2 gloo processes build their shards, trade halo request lists with
torch.distributed, and run the pack -> send/recv -> unpack exchange on CPU tensors with the library's own lists."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


def _build(ib, for_rank=None):
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    m = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12)),
                refinement_regions=[(ib.Ball([0, 0, 0], 0.9), F32(0.24))])
    if for_rank is not None:   # what bench.py does on every rank: ghosts searched in the rank's block range only
        return m, fams, ib.Domain(m, max_partition_size=len(m), hypercube_families=fams, build_partitions=False, build_surfaces=False,
                                  upload=False, for_rank=for_rank)
    return m, fams, ib.Domain(m, hypercube_families=fams, build_partitions=False, upload=False)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import immersedboundary_jl_b200 as ib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, _, g = _build(ib, for_rank=(rank, world))

        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        loc = g.shard(rank, world, all_gather_object=gather)
        info = loc.shard_info
        l2g = info["local_to_global"]
        n_local = info["n_owned"] + info["n_halo"]
        f = lambda gid: np.stack([np.sin(gid * 0.37), gid * 1.0], axis=1).astype(F32)  # a field known by global id
        A = np.zeros((n_local, 2), F32)
        A[: info["n_owned"]] = f(l2g[: info["n_owned"]].astype(np.float64))
        lists = loc.send_lists()
        reqs = []
        recv_bufs = {}
        for peer in range(world):
            s_, r_ = lists[peer]
            if len(s_):
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(A[s_])), peer))
            if len(r_):
                recv_bufs[peer] = torch.zeros((len(r_), 2), dtype=torch.float32)
                reqs.append(dist.irecv(recv_bufs[peer], peer))
        for r in reqs:
            r.wait()
        for peer, buf in recv_bufs.items():
            A[lists[peer][1]] = buf.numpy()
        needed = np.concatenate([lists[p][1] for p in range(world)])
        ok = np.array_equal(A[needed], f(l2g[needed].astype(np.float64)))
        q.put((rank, ok, info["n_owned"], info["owned_start"], sorted(l2g[needed].tolist())))
    finally:
        dist.destroy_process_group()


def test_two_rank_halo_exchange_gloo(ib, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, fams, g = _build(ib)
    assert sum(r[2] for r in res) == len(g) and res[0][3] == 0 and res[1][3] == res[0][2]
    assert all(r[1] for r in res)
    # the requested cells are exactly the reference's 2-deep skirt of the owned range (+ image donors)
    OM = oracle.mesher
    om = OM.Mesh([-2, -2, -2], [4, 4, 4], ("wall", OM.AnalyticSphere([0, 0, 0], 0.5), F32(0.12)),
                 refinement_regions=[(OM.Ball([0, 0, 0], 0.9), F32(0.24))])
    od = oracle.domain.Domain(om, hypercube_families=fams)
    for rank, _, n_owned, start, needed in res:
        image = np.arange(start, start + n_owned, dtype=np.int64)
        part = oracle.domain.build_partition(1, image, od.faces, od.c2f_ptr, od.c2f_idx, od.centers, od.widths, 2)
        skirt = set(np.setdiff1d(part.domain, image).tolist())
        donors = set()
        for bs in od.boundaries.values():
            for b in bs.values():
                own = (b.ghost_indices >= start) & (b.ghost_indices < start + n_owned)
                ptr, idx, _ = b.image_interpolator.to_csr()
                for gi in np.flatnonzero(own):
                    donors.update(b.image_domain[idx[ptr[gi]:ptr[gi + 1]]].tolist())
        donors = {d for d in donors if not (start <= d < start + n_owned)}
        assert set(needed) == skirt | donors


def test_single_rank_shard_is_identity(ib):
    _, _, g = _build(ib)
    loc = g.shard(0, 1)
    assert loc.shard_info["n_owned"] == len(g) and loc.shard_info["n_halo"] == 0
    assert np.array_equal(loc.shard_info["local_to_global"], np.arange(len(g)))
    assert np.array_equal(loc.block_faces(), g.block_faces())
    for name in g.boundaries:
        assert sum(b.nghost for b in loc.boundaries[name].values()) == sum(b.nghost for b in g.boundaries[name].values())


def _shard_all(ib, msh, fams, world):
    """Every rank's shard in one process: threads with a barrier stand in for all_gather_object (ctypes drops the GIL)."""
    import threading
    g = ib.Domain(msh, hypercube_families=fams, build_partitions=False, upload=False)
    bar, slots, out = threading.Barrier(world), [None] * world, [None] * world

    def worker(r):
        def gather(obj):
            slots[r] = obj
            bar.wait()
            res = list(slots)
            bar.wait()
            return res
        out[r] = g.shard(r, world, all_gather_object=gather).shard_info

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


def test_boundary_families_coupled_across_ranks_are_detected(ib):
    """ghost_update_euler applies the families in sequence; when a ghost of one family interpolates from a ghost of
    another family owned by a different rank, the halo rows must be refreshed in between.  A body far from
    the box needs no extra exchange, a body next to a box face does -- and every rank must agree (the exchange is
    collective)."""
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    far = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.12)),
                  refinement_regions=[(ib.Ball([0, 0, 0], 0.9), F32(0.24))])
    close = ib.Mesh([-1, -1, -1], [2, 2, 2], ("wall", ib.Sphere([-0.55, 0, 0], 0.4), F32(0.1)),
                    refinement_regions=[(ib.Ball([-0.55, 0, 0], 0.6), F32(0.1))])
    for world in (2, 4):
        assert all(i["coupled_families"] == set() for i in _shard_all(ib, far, fams, world))
        infos = _shard_all(ib, close, fams, world)
        assert all(i["coupled_families"] == infos[0]["coupled_families"] for i in infos)
        assert ("farfield", "wall") in infos[0]["coupled_families"] or ("wall", "farfield") in infos[0]["coupled_families"]


@pytest.mark.parametrize("world", [2, 3])
def test_rank_restricted_build_gives_the_same_shard(ib, world):
    """`Domain(..., for_rank=(rank, world))` (ibx_domain_build_for_rank: ghosts searched in the rank's block range only,
    no per-cell array of the global mesh is ever allocated) followed by `.shard(rank, world)` must give the tables that
    sharding the full global build gives: cells, ghosts, projections, donors and weights, halo request lists."""
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    pts, tri = ib.synthetic.icosphere(1, 0.5)
    m = ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Stereolitography(pts, tri), F32(0.12)),
                refinement_regions=[(ib.Ball([0, 0, 0], 0.9), F32(0.24))])
    full = ib.Domain(m, hypercube_families=fams, build_partitions=False, build_surfaces=False, upload=False)
    cen, wid = full.cells()
    # the shard's send lists need every rank's requests: build all shards of both kinds first
    a = [full.shard(r, world, all_gather_object=lambda obj: [obj] * world) for r in range(world)]
    b = []
    for r in range(world):
        g = ib.Domain(m, max_partition_size=len(m), hypercube_families=fams, build_partitions=False, build_surfaces=False,
                      upload=False, for_rank=(r, world))
        assert len(g) == len(full)
        gc, gw = g.cells()                                   # computed on request from the block tables
        assert np.array_equal(gc, cen) and np.array_equal(gw, wid)
        b.append(g.shard(r, world, all_gather_object=lambda obj: [obj] * world))
    for r in range(world):
        A, B = a[r], b[r]
        ia, ib_ = A.shard_info, B.shard_info
        assert (ia["n_owned"], ia["n_halo"], ia["owned_start"]) == (ib_["n_owned"], ib_["n_halo"], ib_["owned_start"])
        assert np.array_equal(ia["local_to_global"], ib_["local_to_global"])
        for peer in range(world):
            assert np.array_equal(ia["requests"][peer], ib_["requests"][peer])
        ca, wa = A.cells()
        cb, wb = B.cells()
        assert np.array_equal(ca, cb) and np.array_equal(wa, wb)
        assert np.array_equal(ca, cen[ia["local_to_global"]])
        assert np.array_equal(A.block_faces(), B.block_faces())
        n_ghost = 0
        for name in A.boundaries:
            assert list(A.boundaries[name]) == list(B.boundaries[name])
            for key, x in A.boundaries[name].items():
                y = B.boundaries[name][key]
                n_ghost += len(x.ghost_indices)
                for field in ("ghost_indices", "image_domain", "projections", "normals_host", "image_distances", "ghost_distances",
                              "interp_ptr", "interp_idx", "interp_w"):
                    assert np.array_equal(getattr(x, field), getattr(y, field)), (name, field)
        assert n_ghost > 100
