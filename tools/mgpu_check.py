"""Multi-GPU parity worker (run under torchrun; tests/test_mgpu_gpu.py launches it for 2 / 4 / 8 ranks): the sharded step
(halo exchange over NCCL + ghost update + halo exchange + residual) must reproduce, on every rank's owned cells, the
single-domain result computed on the same GPU -- bit for bit.

Also checked on every case: the overlapped entry point (ibx_step_euler_sharded) and the C5 RANS step (ibx_residual_rans +
ibx_ghost_update_rans with two exchanged arrays).
Cases: `sphere` (body far from the box: the two boundary families do not interact) and `close` (a sphere next to a box
face: wall ghosts interpolate from farfield ghosts owned by other ranks, so ghost_update_euler has to exchange the halo
rows between the two families -- ADVICE r1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import immersedboundary_jl_b200 as ib

F32 = np.float32
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ib.context(local)
fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]


def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


ident = np.zeros(128, np.uint8)
if rank == 0:
    ib._lib.call("ibx_comm_unique_id", ib._lib.ptr(ident))
obj = [ident.tobytes()]
dist.broadcast_object_list(obj, src=0)
ident = np.frombuffer(obj[0], np.uint8).copy()
ib._lib.call("ibx_comm_init", ctx, rank, world, ib._lib.ptr(ident))
fl = ib.Fluid()
a = np.sqrt(1.4 * 283.0 * 288.15)
bcs = [("wall", ib.FlowBC(fl, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)),
       ("farfield", ib.FlowBC(fl, np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)))]


def case(label, msh, expect_coupled):
    g = ib.Domain(msh, hypercube_families=fams, build_partitions=False, upload=False)
    loc = g.shard(rank, world, all_gather_object=gather)
    loc.upload()
    g.upload()
    info = loc.shard_info
    l2g, n_owned = info["local_to_global"], info["n_owned"]
    Qg0 = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(g.cells()[0]))
    # reference: whole domain on this GPU
    Qg = ib.DeviceArray.from_host(Qg0)
    Rg, cg = ib.DeviceArray(len(g), 5, False), ib.DeviceArray(len(g), 1, True)
    for _ in range(2):
        ib.ghost_update_euler(g, fl, Qg, bcs)
        ib.residual_euler(g, fl, Qg, Rg, cg)
    # sharded: only the owned rows are initialised; halo rows arrive through the exchange
    Ql0 = np.zeros((len(loc), 5), F32)
    Ql0[:n_owned] = Qg0[l2g[:n_owned]]
    Ql = ib.DeviceArray.from_host(Ql0)
    Rl, cl = ib.DeviceArray(len(loc), 5, False), ib.DeviceArray(len(loc), 1, True)
    need = np.concatenate([v for v in info["requests"].values()])
    g2l = {int(gid): i for i, gid in enumerate(l2g)}
    need_local = np.array([g2l[int(x)] for x in need], dtype=np.int64)
    loc.halo_exchange(Ql)
    ib.synchronize()
    got = Ql.to_host()
    ok0 = np.array_equal(got[need_local], Qg0[need])
    for _ in range(2):  # twice: the second pass sees ghost values updated (and exchanged) by the first
        loc.halo_exchange(Ql)
        ib.ghost_update_euler(loc, fl, Ql, bcs)
        loc.halo_begin(Ql)      # completed inside residual_euler (overlapped with compute on the owned rows)
        ib.residual_euler(loc, fl, Ql, Rl, cl)
    # the same two steps through the overlapped entry point (phase split: exchange + ghost update hidden behind the blocks
    # that read neither a ghost nor a halo cell)
    Qf = ib.DeviceArray.from_host(Ql0)
    Rf, cf2 = ib.DeviceArray(len(loc), 5, False), ib.DeviceArray(len(loc), 1, True)
    for _ in range(2):
        ib.step_euler_sharded(loc, fl, bcs, Qf, Rf, cf2)
    ib.synchronize()
    okF = (np.array_equal(Rf.to_host()[:n_owned], Rl.to_host()[:n_owned]) and np.array_equal(cf2.to_host()[:n_owned], cl.to_host()[:n_owned])
           and np.array_equal(Qf.to_host()[:n_owned], Ql.to_host()[:n_owned]))
    # configuration C5 on the same shard: RANS residual + ghost update of the transported variable (two exchanged arrays)
    nu3 = 4.5e-5
    rbc = [("wall", 0.0), ("farfield", nu3)]
    qg0 = (Qg0[:, 0] * F32(nu3) * (1 + F32(0.2) * np.sin(g.cells()[0][:, 0]).astype(F32))).astype(F32)
    Qg2, qg = ib.DeviceArray.from_host(Qg0), ib.DeviceArray.from_host(qg0)
    Rg2, RRg, cg2 = ib.DeviceArray(len(g), 5, False), ib.DeviceArray(len(g), 1, True), ib.DeviceArray(len(g), 1, True)
    ql0 = np.zeros(len(loc), F32)
    ql0[:n_owned] = qg0[l2g[:n_owned]]
    Ql2, ql = ib.DeviceArray.from_host(Ql0), ib.DeviceArray.from_host(ql0)
    Rl2, RRl, cl2 = ib.DeviceArray(len(loc), 5, False), ib.DeviceArray(len(loc), 1, True), ib.DeviceArray(len(loc), 1, True)
    for _ in range(2):
        ib.ghost_update_euler(g, fl, Qg2, bcs)
        ib.ghost_update_rans(g, Qg2, qg, rbc)
        ib.residual_rans(g, fl, Qg2, qg, Rg2, RRg, cg2)
        loc.halo_exchange(Ql2)
        loc.halo_exchange(ql)
        ib.ghost_update_euler(loc, fl, Ql2, bcs)
        loc.halo_exchange(Ql2)      # R = qR / rho at a donor that is itself a ghost cell uses its NEW density: refresh Q first
        ib.ghost_update_rans(loc, Ql2, ql, rbc)
        loc.halo_exchange(ql)
        ib.residual_rans(loc, fl, Ql2, ql, Rl2, RRl, cl2)
    own = l2g[:n_owned]
    okC5 = (np.array_equal(Rl2.to_host()[:n_owned], Rg2.to_host()[own]) and np.array_equal(RRl.to_host()[:n_owned], RRg.to_host()[own])
            and np.array_equal(ql.to_host()[:n_owned], qg.to_host()[own]) and np.array_equal(cl2.to_host()[:n_owned], cg2.to_host()[own]))
    if not okC5:
        for lab, a, b in (("R5", Rl2.to_host()[:n_owned], Rg2.to_host()[own]), ("RR", RRl.to_host()[:n_owned], RRg.to_host()[own]),
                          ("qR", ql.to_host()[:n_owned], qg.to_host()[own]), ("cfl", cl2.to_host()[:n_owned], cg2.to_host()[own])):
            bad = np.flatnonzero((a != b).reshape(len(a), -1).any(axis=1))
            print(f"[{label}] rank {rank}: C5 {lab}: {len(bad)} rows differ, max |diff| {np.abs(a - b).max():.3e} of scale {np.abs(b).max():.3e}; "
                  f"first rows {bad[:6]}", flush=True)
    okR = np.array_equal(Rl.to_host()[:n_owned], Rg.to_host()[own])
    okc = np.array_equal(cl.to_host()[:n_owned], cg.to_host()[own])
    okQ = np.array_equal(Ql.to_host()[:n_owned], Qg.to_host()[own])
    dR = np.abs(Rl.to_host()[:n_owned] - Rg.to_host()[own]).max(axis=1)
    coupled = sorted(info.get("coupled_families", ()))
    print(f"[{label}] rank {rank}: owned {n_owned} halo {info['n_halo']} exchanged {len(need)} coupled {coupled} | first exchange {ok0} "
          f"Q {okQ} R {okR} cfl {okc} overlapped-step {okF} C5 {okC5} (cells with R mismatch {(dR > 0).sum()})", flush=True)
    ok = ok0 and okR and okc and okQ and okF and okC5 and (bool(coupled) == expect_coupled)
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.item() == 1


ok = case("sphere", ib.Mesh([-2, -2, -2], [4, 4, 4], ("wall", ib.Sphere([0, 0, 0], 0.5), F32(0.06)),
                            refinement_regions=[(ib.Ball([0, 0, 0], 0.9), F32(0.12))]), False)
ok &= case("close", ib.Mesh([-1, -1, -1], [2, 2, 2], ("wall", ib.Sphere([-0.55, 0, 0], 0.4), F32(0.05)),
                           refinement_regions=[(ib.Ball([-0.55, 0, 0], 0.6), F32(0.05))]), True)
ib._lib.call("ibx_comm_finalize", ctx)
dist.destroy_process_group()
if rank == 0:
    print("MGPU PARITY", "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
