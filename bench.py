#!/usr/bin/env python
"""Benchmark of the hot path: cell-updates/s of (IB ghost update of every boundary + full Euler residual).

    python bench.py --gpus N --steps K --warmup W            # our arm (libibx, sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the restated reference path (compiled C + OpenMP, all host cores)
    python bench.py --workload c5 [--gpus N]                 # secondary workload: swept-wing RANS residual (configs[4])

Workload (BASELINE.json configs[3], SURVEY.md 8d "C4"): 3-D sphere octree, box (-16)^3..16^3, block_size 8,
growth_ratio 2, wall = icosphere STL (6 subdivisions, 81 920 triangles, radius 0.5), finest cell width 32/2^10/8 inside Ball(0, r); r = 0.75
gives 50.2 M cells on one GPU and is enlarged so that the cell count grows with the number of GPUs (weak
scaling, ~50 M cells per GPU).  Inputs: the smooth synthetic state of SURVEY.md 8(d), seed 12345.
One step = [halo exchange +] ghost update (wall, farfield) [+ halo exchange] + residual on every rank, through
ibx_step_euler / ibx_step_euler_sharded (exchanges and ghost update run on a second stream under ghost-free work).
The JSON line also carries `fast_mode` (the price of the bit-exact arithmetic), `halo` (phase timings at N > 1), `e2e`
(host buffers in and out, two evaluations in flight) and `cpu_baseline`.
The working set (state 1 GB + residual 1.2 GB + scratch 1.2 GB per 50 M cells) is far larger than the 126 MB L2,
so no explicit L2 flush is needed between timed iterations.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
F32 = np.float32

H_FINE = 32.0 / 2 ** 10 / 8 * 1.01       # requested cell size: just above the level-10 cell width
CELLS_PER_GPU = 50_000_000
B_ALG = 4 * (2 * 5 + 1)                  # 44 B per cell-update: read Q, write R, write cfl
B_GHOST = 4 * (5 + 5) + 8 * (4 + 4) + 4 * (3 + 1)   # 120 B per ghost


STL_SUBDIV = 6                           # icosphere subdivisions of the wall STL (81 920 triangles)
WORKLOAD = ("C4: 3-D sphere octree Euler residual (MUSCL + JST sensor + HLL) + IB ghost update (wall, farfield); wall = icosphere "
            "STL, 81 920 triangles, radius 0.5")


def build_mesh(ib, radius, h=H_FINE, analytic=False):
    """C4 recipe: the wall is an icosphere STL refined by the reference's own `refine_to_length` rule inside
    `Mesh`; `analytic=True` swaps in the exact sphere (an extension of this library, kept for quick smoke runs)."""
    if analytic:
        surf = ib.Sphere([0, 0, 0], 0.5)
    else:
        pts, tri = ib.synthetic.icosphere(STL_SUBDIV, 0.5)
        surf = ib.Stereolitography(pts, tri)
    return ib.Mesh([-16, -16, -16], [32, 32, 32], ("wall", surf, F32(h)), refinement_regions=[(ib.Ball([0, 0, 0], radius), F32(h))])


def radius_for(ib, target_cells, h=H_FINE):
    """Smallest refinement-ball radius (on a 1/64 grid) whose mesh has at least `target_cells` cells.  The block list
    around the ball does not depend on how the sphere surface is represented (the ball contains it), so the search
    runs on the analytic sphere."""
    if target_cells <= 50_300_000 and h == H_FINE:
        return 0.75
    lo, hi = 0.5, 4.0
    for _ in range(12):
        mid = round((lo + hi) / 2 * 64) / 64
        if mid in (lo, hi):
            break
        if len(build_mesh(ib, mid, h, analytic=True)) >= target_cells:
            hi = mid
        else:
            lo = mid
    return hi


class ClockSampler:
    """SM clock / power / throttle reasons polled through NVML every ~5 ms from a thread, started BEFORE the warm-up so
    that a sub-100 ms timed region is still covered (B200_PROFILING.md's clocks line; nvidia-smi -lms cannot go that fast).
    Only the samples between mark_begin() and mark_end() enter the median."""

    def __init__(self, device):
        self.rows, self.device, self.stop_flag, self.thread = [], device, False, None
        self.t0 = self.t1 = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[device])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0, nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                pass
            time.sleep(0.004)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30)] or self.rows
        sm = [r[1] for r in rows]
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = [n for n, b in bits.items() if any(r[3] & b for r in rows)]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w_max": max(r[2] for r in rows) if rows else None,
                "samples": len(sm), "samples_total": len(self.rows), "reasons": reasons, "how": "NVML polled every ~5 ms during the timed region"}


def flux_scaled_error(dom_cells_widths, Q, R, R_ref, gamma=1.4, Rgas=283.0):
    """|r - r_ref| / max(|r_ref|, sum_faces |F| / dx) per cell and variable, with the face-flux
    magnitude estimated from the cell state (|F| of the physical flux in each direction, two faces per direction)."""
    Q = Q.astype(np.float64)
    nd = Q.shape[1] - 2
    rho, E = Q[:, 0], Q[:, 1]
    u = Q[:, 2:] / rho[:, None]
    p = (gamma - 1.0) * (E - 0.5 * rho * (u ** 2).sum(axis=1))
    S = np.zeros_like(Q)
    for d in range(nd):
        F = Q * u[:, d:d + 1]
        F[:, 1] += p * u[:, d]
        F[:, 2 + d] += p
        S += 2.0 * np.abs(F) / dom_cells_widths[:, d:d + 1].astype(np.float64)
    return np.abs(R.astype(np.float64) - R_ref) / np.maximum(np.abs(R_ref.astype(np.float64)), S)


def measured_peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------- CPU arm
def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def cpu_reference(level, radius, steps, warmup, stl=None):
    """The restated reference path on all host cores (oracle/cpu_ref.c: compiled C + OpenMP, one task per partition like
    ThreadTools.tmap, gather -> array-at-a-time operators with temporaries -> scatter; bit-identical to the NumPy oracle,
    tests/test_oracle_cpu_ref.py) on the bench's own sphere-octree recipe at octree level `level`.

    The partition / boundary tables and the input state come from tools/dump_cpu_tables.py, run as a SEPARATE, untimed
    process: this process never loads libibx.so (the reference arm must not depend on the product)."""
    import shutil
    import tempfile
    from oracle import cfd as ocfd, cpu_ref
    cores = os.cpu_count() or 1
    cpu_ref.build()
    tmp = tempfile.mkdtemp(prefix="ibx_cpu_tables_")
    try:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "dump_cpu_tables.py"), tmp, str(level), str(radius), str(400_000)]
        if stl is not None:
            cmd += ["stl", str(stl)]
        # torchrun exports OMP_NUM_THREADS=1: the (untimed) builder process gets all cores back
        subprocess.run(cmd, check=True, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS=str(cores)))
        ref, Q, meta = cpu_ref.CpuRef.from_dump(tmp)
        n, nparts = meta["ncells"], len(meta["parts"])
        fl = ocfd.Fluid()
        a = np.sqrt(1.4 * 283.0 * 288.15)
        Pinf = np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)
        bcs = [("wall", ocfd.FlowBC(fl, Pinf[:3] * np.array([1, 1, 0], F32), normal_flow=True)), ("farfield", ocfd.FlowBC(fl, Pinf))]
        R, cf = np.zeros((n, 5), F32, order="F"), np.zeros(n, F32)

        def step():
            ref.ghost_update(fl, Q, bcs, cores)
            ref.residual(fl, Q, R, cf, cores)

        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert "libibx.so" not in open("/proc/self/maps").read() or "immersedboundary_jl_b200" in sys.modules
    return {"value": n * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"compiled C + OpenMP restatement of the reference operators (oracle/cpu_ref.c; Julia cannot run here), the "
                      f"bench's sphere-octree recipe ({meta['surface']}) at octree level {level}, refinement ball {radius} -> {n} cells "
                      f"in {nparts} partitions, {steps} evaluations (ghost update + residual) after {warmup} warm-up, {cores} threads "
                      f"(one task per partition, like tmap); tables built by an untimed helper process",
            "ms_per_step": dt / steps * 1e3, "cells": n, "level": level}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the benched mesh itself (level 10, 50.2 M cells) when the host has the memory for its tables (~45 GB peak in the helper
    # process), else one octree level coarser at a ball radius giving ~20 M cells
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    r = None
    stl = None if args.analytic_sphere else STL_SUBDIV
    if args.level != 10:
        r = cpu_reference(args.level, 0.75, steps, warmup, stl)           # smoke runs / unit test of the contract line
    elif _mem_available_gb() > 110:
        try:
            r = cpu_reference(10, 0.75, steps, warmup, stl)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"[bench] full-size CPU arm failed ({e}); falling back to the level-9 sample\n")
    if r is None:
        r = cpu_reference(9, 1.0, steps, warmup, stl)
    line = {"impl": "reference", "metric": "cell-updates/s (Euler residual+IB)", "value": r["value"], "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + " -- CPU arm", "cells": r["cells"], "finest_level": r["level"], "block_size": 8, "nv": 5},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import immersedboundary_jl_b200 as ib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # torchrun exports OMP_NUM_THREADS=1; the host-side builder is OpenMP code: give every rank its share of the cores
        try:
            C.CDLL("libgomp.so.1").omp_set_num_threads(max(1, (os.cpu_count() or 8) // world))
        except OSError:
            pass
    ctx = ib.context(local_rank)
    t_setup = time.perf_counter()
    cells_target = args.cells if args.cells else CELLS_PER_GPU * world
    h = 32.0 / 2 ** args.level / 8 * 1.01
    radius = radius_for(ib, cells_target, h) if not args.radius else args.radius
    msh = build_mesh(ib, radius, h, analytic=args.analytic_sphere)
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    gdom = ib.Domain(msh, max_partition_size=len(msh), hypercube_families=fams, build_partitions=False,
                     build_surfaces=False, upload=False, for_rank=(rank, world) if world > 1 else None)
    n_global = len(gdom)
    if world > 1:
        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        dom = gdom.shard(rank, world, all_gather_object=gather)
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            ib._lib.call("ibx_comm_unique_id", ib._lib.ptr(ident))
        obj = [ident.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        ident = np.frombuffer(obj[0], np.uint8).copy()
        ib._lib.call("ibx_comm_init", ctx, rank, world, ib._lib.ptr(ident))
        n_owned = dom.shard_info["n_owned"]
        del gdom
        centers = dom.cells()[0]
    else:
        dom, n_owned = gdom, n_global
        centers = dom.cells()[0]
    dom.upload()
    n_local = len(dom)
    fluid = ib.Fluid()
    a_inf = math.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a_inf, 0.0, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fluid, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)), ("farfield", ib.FlowBC(fluid, Pinf))]
    Q_host = ib.pinned_empty((n_local, 5))
    Q_host[...] = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(centers))
    del centers
    Q = ib.DeviceArray(n_local, 5, False).upload(Q_host)
    R, cfl = ib.DeviceArray(n_local, 5, False), ib.DeviceArray(n_local, 1, True)
    n_ghost = sum(b.nghost for bs in dom.boundaries.values() for b in bs.values())
    n_tied = {name: int(sum(b.n_tied for b in bs.values())) for name, bs in dom.boundaries.items()} if world == 1 else None
    setup_s = time.perf_counter() - t_setup

    def step():
        if world > 1:
            # exchange -> ghost update -> exchange on the halo stream, hidden behind the residual of the blocks that read
            # neither a ghost nor a halo cell (ibx_step_euler_sharded)
            ib.step_euler_sharded(dom, fluid, bcs, Q, R, cfl)
        else:
            ib.step_euler(dom, fluid, bcs, Q, R, cfl)   # ghost updates on a second stream under the ghost-free blocks' residual

    def barrier():
        ib.synchronize()
        if world > 1:
            dist.barrier()
        ib.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ib.launch_count()
    sampler.mark_begin()
    ib._lib.call("ibx_timer_start", ctx)
    for _ in range(args.steps):
        step()
    ms = C.c_float()
    ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
    sampler.mark_end()
    barrier()
    launches = ib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms.value
    if world > 1:
        import torch
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_global * args.steps / (ms_total * 1e-3)

    # where a sharded step spends its time: each phase timed on its own (CUDA events, max over ranks), and the exchange volume
    halo = None
    if world > 1:
        import torch
        sc, rc_ = np.zeros(world, np.int64), np.zeros(world, np.int64)
        ib._lib.call("ibx_halo_sizes", dom._h, world, ib._lib.ptr(sc), ib._lib.ptr(rc_))

        def timed(fn, reps=20):
            for _ in range(2):      # untimed: the first whole-domain call after the phased steps grows a scratch buffer
                fn()
            barrier()
            ib._lib.call("ibx_timer_start", ctx)
            for _ in range(reps):
                fn()
            ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
            tt = torch.tensor([ms.value / reps], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        ms_x = timed(lambda: dom.halo_exchange(Q))
        ms_g = timed(lambda: ib.ghost_update_euler(dom, fluid, Q, bcs))
        ms_r = timed(lambda: ib.residual_euler(dom, fluid, Q, R, cfl))
        vol = torch.tensor([float(sc.sum()), float(rc_.sum()), float((sc > 0).sum())], device="cuda")
        vmax = vol.clone()
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        ne, nl = C.c_int64(), C.c_int64()
        ib._lib.call("ibx_shard_phase_info", dom._h, C.byref(ne), C.byref(nl))
        halo = {"exchanges_per_step": 2, "owned_blocks_early_phase_rank0": ne.value, "owned_blocks_late_phase_rank0": nl.value, "collective": "grouped ncclSend/ncclRecv (pack kernel -> NCCL -> unpack kernel)",
                "rows_sent_max_rank": int(vmax[0].item()), "rows_received_max_rank": int(vmax[1].item()),
                "bytes_sent_per_exchange_max_rank": int(vmax[0].item()) * 20, "peers_max_rank": int(vmax[2].item()),
                "ms_exchange_alone": ms_x, "ms_ghost_update_alone": ms_g, "ms_residual_alone": ms_r,
                "ms_step": ms_step, "ms_exposed_communication": ms_step - ms_g - ms_r,
                "overlap": "ibx_step_euler_sharded: exchange 1 + ghost update + exchange 2 on the high-priority halo stream under the residual "
                           "of the blocks that read neither a ghost nor a halo cell",
                "coupled_families": sorted(dom.shard_info.get("coupled_families", ())),
                "limiter": "latency of the two exchanges (3 dependent launches + NCCL each), not bandwidth: "
                           f"{int(vmax[0].item()) * 20 / 1e6:.1f} MB per exchange"}

    # dominant kernel alone (CUDA events on the library's compute stream) for the roofline entry
    reps = max(3, min(args.steps, 10))
    for _ in range(2):      # untimed: the first whole-domain call after the phased steps grows a scratch buffer
        ib.residual_euler(dom, fluid, Q, R, cfl)
    ib._lib.call("ibx_timer_start", ctx)
    for _ in range(reps):
        ib.residual_euler(dom, fluid, Q, R, cfl)
    ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
    ms_res = ms.value / reps
    peak, peak_kind = measured_peak_gbs()
    alg_bytes = n_owned * B_ALG
    # DRAM bytes of one call from the committed ncu launch list (only valid for the mesh it was captured on)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic_c4.json")))
        if world == 1 and tr["cells"] == n_global:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    achieved = alg_bytes / (ms_res * 1e-3) / 1e9

    # the price of the bit-exact policy: option "arithmetic" = 1 (Float32 flux, FMA contraction, approximate reciprocals)
    # timed on the same mesh and state, with its error against the exact residual under BOTH normalisations -- SURVEY.md
    # section 7 (scaled by the face fluxes) and the stricter one of tests/test_fused_gpu.py (scaled by the residual)
    fast = None
    if world == 1 and not args.no_fast_mode:
        with ib.options(arithmetic=1):
            for _ in range(2):
                step()
            ib._lib.call("ibx_timer_start", ctx)
            for _ in range(args.steps):
                step()
            ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
            ms_fast_step = ms.value / args.steps
            ib._lib.call("ibx_timer_start", ctx)
            for _ in range(reps):
                ib.residual_euler(dom, fluid, Q, R, cfl)
            ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
            ms_fast_res = ms.value / reps
            R_fast = R.to_host()
        ib.residual_euler(dom, fluid, Q, R, cfl)      # the exact arithmetic on the SAME state (the ghost updates above moved Q)
        R_exact = R.to_host()
        Qh_state = Q.to_host()
        e_flux = flux_scaled_error(dom.cells()[1], Qh_state, R_fast, R_exact)
        scale = np.abs(R_exact).max(axis=0)
        e_res = np.abs(R_fast - R_exact) / np.maximum(np.abs(R_exact), 1e-3 * scale)
        fast = {"option": "arithmetic = 1 (ibx_set_option): Float32 HLL + Green-Gauss, FMA contraction, approximate reciprocals",
                "ms_per_step": ms_fast_step, "value": n_global / (ms_fast_step * 1e-3), "ms_per_residual": ms_fast_res,
                "roofline_frac": alg_bytes / (ms_fast_res * 1e-3) / 1e9 / peak,
                "err_flux_scaled": {"max": float(e_flux.max()), "p99_9": float(np.quantile(e_flux[:, 0], 0.999)),
                                    "definition": "|r - r_exact| / max(|r_exact|, sum_faces |F| / dx)  (SURVEY.md section 7)"},
                "err_residual_scaled": {"max": float(e_res.max()), "median": float(np.median(e_res[:, 0])),
                                        "definition": "|r - r_exact| / max(|r_exact|, 1e-3 max|r_exact|)  (tests/test_fused_gpu.py)"},
                "north_star_1e-5": {"flux_scaled": bool(e_flux.max() < 1e-5), "residual_scaled": bool(e_res.max() < 1e-5)},
                "note": "`value` above is the exact arithmetic (bit-identical to the oracle); this block prices the alternative"}
        del R_exact, R_fast, Qh_state, e_flux, e_res

    # end-to-end through the C ABI with HOST buffers (pinned): every step copies ITS state host->device, runs ghost
    # update + residual, and copies R and cfl device->host.  Consecutive steps are independent evaluations (two sets
    # of host buffers, as for finite-difference JVP probes), enqueued on alternating slots so that the upload of one
    # overlaps the download of the other (PCIe is full duplex); the serial single-call form is timed beside it.
    e2e = None
    if world == 1:
        del R, cfl
        Qh = [Q_host, ib.pinned_empty((n_local, 5))]
        Qh[1][...] = Q_host
        Rh = [ib.pinned_empty((n_local, 5)) for _ in range(2)]
        ch = [ib.pinned_empty((n_local,)) for _ in range(2)]
        for _ in range(2):
            ib.euler_step_host(dom, fluid, bcs, Qh[0], Rh[0], ch[0])
        k = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            ib.euler_step_host(dom, fluid, bcs, Qh[0], Rh[0], ch[0])
        dt_serial = time.perf_counter() - t0
        kp = max(4, min(args.steps, 12))
        for i in range(2):
            ib.euler_step_host_begin(dom, fluid, bcs, Qh[i], Rh[i], ch[i], i)
        for i in range(2):
            ib.euler_step_host_end(i)
        t0 = time.perf_counter()
        for i in range(kp):
            sl = i % 2
            ib.euler_step_host_end(sl)          # results of step i - 2 are in Rh[sl], ch[sl]
            ib.euler_step_host_begin(dom, fluid, bcs, Qh[sl], Rh[sl], ch[sl], sl)
        for i in range(2):
            ib.euler_step_host_end(i)
        dt = time.perf_counter() - t0
        e2e = {"value": n_global * kp / dt, "unit": "cell-updates/s", "h2d_bytes_per_step": int(Q_host.nbytes),
               "d2h_bytes_per_step": int(Rh[0].nbytes + ch[0].nbytes), "ms_per_step": dt / kp * 1e3, "steps": kp,
               "api": "ibx_euler_step_host_begin/_end (C ABI, pinned host buffers, two slots in flight)",
               "serial_single_call": {"value": n_global * k / dt_serial, "ms_per_step": dt_serial / k * 1e3, "steps": k,
                                      "api": "ibx_euler_step_host"}}
    else:
        # every rank: H2D of its shard's state (owned + halo rows) from pinned memory, the sharded step (halo exchange,
        # ghost update, halo exchange, residual), D2H of its residual and CFL arrays; max over ranks.  Two independent
        # evaluations in flight on alternating slots, like the single-GPU arm: the upload of one overlaps the compute and the
        # download of the other (ibx_array_upload_async / _download_async / ibx_download_fence / _wait).
        del R, cfl
        Qd = [Q, ib.DeviceArray(n_local, 5, False)]
        Rd = [ib.DeviceArray(n_local, 5, False) for _ in range(2)]
        cd = [ib.DeviceArray(n_local, 1, True) for _ in range(2)]
        Qh = [Q_host, ib.pinned_empty((n_local, 5))]
        Qh[1][...] = Q_host
        Rh = [ib.pinned_empty((n_local, 5)) for _ in range(2)]
        ch = [ib.pinned_empty((n_local,)) for _ in range(2)]

        def begin(sl):
            ib._lib.call("ibx_array_upload_async", ctx, Qd[sl].h, ib._lib.ptr(Qh[sl]))
            ib.step_euler_sharded(dom, fluid, bcs, Qd[sl], Rd[sl], cd[sl])
            ib._lib.call("ibx_array_download_async", ctx, Rd[sl].h, ib._lib.ptr(Rh[sl]))
            ib._lib.call("ibx_array_download_async", ctx, cd[sl].h, ib._lib.ptr(ch[sl]))
            ib._lib.call("ibx_download_fence", ctx, sl)

        def end(sl):
            ib._lib.call("ibx_download_wait", ctx, sl)

        for sl in range(2):
            begin(sl)
        for sl in range(2):
            end(sl)
        k = max(4, min(args.steps, 12))
        barrier()
        t0 = time.perf_counter()
        for i in range(k):
            sl = i % 2
            end(sl)
            begin(sl)
        for sl in range(2):
            end(sl)
        barrier()
        import torch
        t = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        hb = torch.tensor([float(Q_host.nbytes), float(Rh[0].nbytes + ch[0].nbytes)], device="cuda")
        dist.all_reduce(hb)
        e2e = {"value": n_global * k / dt, "unit": "cell-updates/s", "h2d_bytes_per_step": int(hb[0].item()),
               "d2h_bytes_per_step": int(hb[1].item()), "ms_per_step": dt / k * 1e3, "steps": k,
               "host_gbs_aggregate": (hb[0].item() + hb[1].item()) * k / dt / 1e9,
               "api": "ibx_array_upload_async + sharded step + ibx_array_download_async on every rank, two slots in flight "
                      "(C ABI, pinned host buffers)"}
        R, cfl = Rd[0], cd[0]

    if world > 1:
        ib._lib.call("ibx_comm_finalize", ctx)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample: the same recipe one octree level coarser (10.3 M cells), ~10-30 s of CPU work
        r = cpu_reference(9, 0.75, 8, 1, None if args.analytic_sphere else STL_SUBDIV)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": "cell-updates/s (Euler residual+IB)", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if not args.analytic_sphere else WORKLOAD.split(";")[0] + "; wall = analytic sphere (smoke option)",
                   "ghosts_with_tied_donor_candidates": n_tied,
                   "cells": n_global, "cells_per_gpu": n_owned, "ghost_cells_rank0": n_ghost, "blocks": msh.nblocks,
                   "refinement_ball_radius": radius, "finest_level": args.level, "block_size": 8, "nv": 5, "partition": "contiguous block ranges, one per GPU",
                   "l2": "working set >> 126 MB L2, no flush needed", "setup_s": round(setup_s, 1)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "halo": halo,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind,
                     "kernel": "ibx_residual_euler (k_prim + sensor kernels + general-face pass + k_march_flux)",
                     "ms_per_launch": ms_res, "algorithmic_bytes_per_cell": B_ALG,
                     "whole_step_achieved": (n_owned * B_ALG + n_ghost * B_GHOST) / (ms_step * 1e-3) / 1e9},
        "cpu_baseline": cpu, "fast_mode": fast,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- C5 (secondary workload)
def build_wing_mesh(ib, level, margin):
    """C5 recipe: same box as C4, wall = procedurally generated swept wing (NACA 0012 section, span 4, sweep
    30 degrees) as an STL refined by the reference's rule, finest cells inside a box `margin` around the wing."""
    pts, tri = ib.synthetic.swept_wing()
    h = F32(32.0 / 2 ** level / 8 * 1.01)
    lo = pts.min(axis=0) - margin
    w = pts.max(axis=0) - pts.min(axis=0) + 2 * margin
    return ib.Mesh([-16, -16, -16], [32, 32, 32], ("wall", ib.Stereolitography(pts, tri), h),
                   refinement_regions=[(ib.Box(lo.tolist(), w.tolist()), h)])


def run_c5(args):
    """`--workload c5`: 3-D wing RANS residual (Euler part + viscous fluxes with the Wray-Agarwal eddy viscosity + transported-R
    residual, ibx_residual_rans) + IB ghost updates of the mean flow and of R, ~25 M cells per GPU (200 M on 8)."""
    import immersedboundary_jl_b200 as ib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        try:
            C.CDLL("libgomp.so.1").omp_set_num_threads(max(1, (os.cpu_count() or 8) // world))
        except OSError:
            pass
    ctx = ib.context(local_rank)
    t_setup = time.perf_counter()
    target = args.cells if args.cells else 25_000_000 * world
    level = args.level if args.level != 10 or target > 60_000_000 else 9
    # smallest margin (1/64 grid) whose mesh reaches the target; the two standard sizes are tabulated (the search builds
    # the mesh, STL refinement included, six times)
    known = {25_000_000: (9, 0.03125), 200_000_000: (10, 0.140625)}       # -> 29 952 000 and 195 217 408 cells
    lo_m, hi_m = 0.0, 1.0
    if target in known and args.level == 10:
        level, hi_m = known[target]
    for _ in range(0 if target in known and args.level == 10 else 6):
        mid = round((lo_m + hi_m) / 2 * 64) / 64
        if mid in (lo_m, hi_m):
            break
        if len(build_wing_mesh(ib, level, mid)) >= target:
            hi_m = mid
        else:
            lo_m = mid
    msh = build_wing_mesh(ib, level, hi_m)
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    gdom = ib.Domain(msh, max_partition_size=len(msh), hypercube_families=fams, build_partitions=False, build_surfaces=False,
                     upload=False, for_rank=(rank, world) if world > 1 else None)
    n_global = len(gdom)
    if world > 1:
        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        dom = gdom.shard(rank, world, all_gather_object=gather)
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            ib._lib.call("ibx_comm_unique_id", ib._lib.ptr(ident))
        obj = [ident.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        ib._lib.call("ibx_comm_init", ctx, rank, world, ib._lib.ptr(np.frombuffer(obj[0], np.uint8).copy()))
        n_owned = dom.shard_info["n_owned"]
        del gdom
    else:
        dom, n_owned = gdom, n_global
    centers = dom.cells()[0]
    dom.upload()
    n_local = len(dom)
    fluid = ib.Fluid()
    a_inf = math.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a_inf, 0.0, 0.0], F32)
    bcs = [("wall", ib.FlowBC(fluid, np.array([101325.0, 288.15, 0.0], F32), normal_flow=True)), ("farfield", ib.FlowBC(fluid, Pinf))]
    nu_inf = 1.79e-5 / (101325.0 / (283.0 * 288.15))
    rbc = [("wall", 0.0), ("farfield", 3.0 * nu_inf)]              # `R_inf = 3 nu`, `R = 0` at walls (src/turbulence.jl:203)
    Q_host = ib.pinned_empty((n_local, 5))
    Q_host[...] = ib.synthetic.primitive2state_host(ib.synthetic.euler_state(centers))
    q_host = ib.pinned_empty((n_local,))
    q_host[...] = Q_host[:, 0] * F32(3.0 * nu_inf) * (1 + F32(0.2) * np.sin(centers[:, 0]).astype(F32))
    del centers
    Q, qR = ib.DeviceArray(n_local, 5, False).upload(Q_host), ib.DeviceArray(n_local, 1, True).upload(q_host)
    R, RR, cfl = ib.DeviceArray(n_local, 5, False), ib.DeviceArray(n_local, 1, True), ib.DeviceArray(n_local, 1, True)
    n_ghost = sum(b.nghost for bs in dom.boundaries.values() for b in bs.values())
    setup_s = time.perf_counter() - t_setup

    def step():
        if world > 1:
            dom.halo_exchange(Q)
            dom.halo_exchange(qR)
        ib.ghost_update_euler(dom, fluid, Q, bcs)
        if world > 1:
            dom.halo_exchange(Q)      # ghost_update_rans divides by the NEW ghost densities, also at donors owned by other ranks
        ib.ghost_update_rans(dom, Q, qR, rbc)
        if world > 1:
            dom.halo_exchange(qR)
        ib.residual_rans(dom, fluid, Q, qR, R, RR, cfl)

    def barrier():
        ib.synchronize()
        if world > 1:
            dist.barrier()
        ib.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ib.launch_count()
    ms = C.c_float()
    sampler.mark_begin()
    ib._lib.call("ibx_timer_start", ctx)
    for _ in range(args.steps):
        step()
    ib._lib.call("ibx_timer_stop", ctx, C.byref(ms))
    sampler.mark_end()
    barrier()
    launches = ib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms.value
    if world > 1:
        import torch
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    # end to end with host buffers: upload Q and qR, step, download R, RR, cfl (serial form)
    R_host, RR_host, c_host = ib.pinned_empty((n_local, 5)), ib.pinned_empty((n_local,)), ib.pinned_empty((n_local,))
    k = max(2, min(args.steps, 4))
    barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        Q.upload(Q_host)
        qR.upload(q_host)
        step()
        for a, hbuf in ((R, R_host), (RR, RR_host), (cfl, c_host)):
            ib._lib.call("ibx_array_download", ctx, a.h, ib._lib.ptr(hbuf))
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        import torch
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        ib._lib.call("ibx_comm_finalize", ctx)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak_gbs()
    b_alg = 4 * (2 * 6 + 1)                         # 52 B per cell-update (nv = 6)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_c5()
    line = {
        "metric": "cell-updates/s (RANS residual+IB)", "value": n_global * args.steps / (ms_total * 1e-3), "unit": "cell-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5: 3-D swept-wing RANS residual (Euler part: MUSCL + JST + HLL; viscous fluxes with the Wray-Agarwal eddy "
                               "viscosity; transported-R residual) + IB ghost updates (mean flow and R); wall = procedural NACA 0012 wing STL, "
                               "span 4, sweep 30 deg", "cells": n_global, "cells_per_gpu": n_owned, "ghost_cells_rank0": n_ghost,
                   "blocks": msh.nblocks, "finest_level": level, "refinement_margin": hi_m, "block_size": 8, "nv": 6,
                   "l2": "working set >> 126 MB L2, no flush needed", "setup_s": round(setup_s, 1)},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": n_global * k / dt, "unit": "cell-updates/s", "h2d_bytes_per_step": int(Q_host.nbytes + q_host.nbytes) * world,
                "d2h_bytes_per_step": int(R_host.nbytes + RR_host.nbytes + c_host.nbytes) * world, "ms_per_step": dt / k * 1e3, "steps": k,
                "api": "ibx_array_upload + sharded RANS step + ibx_array_download (C ABI, pinned host buffers)"},
        "roofline": {"bound": "hbm", "achieved": n_owned * b_alg / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": n_owned * b_alg / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_kind": peak_kind,
                     "kernel": "whole step: ibx_residual_euler kernels + k_rans_state / _grad / _source / _flux + ghost updates",
                     "algorithmic_bytes_per_cell": b_alg},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_reference_c5():
    """The restated RANS composition (NumPy oracle; no compiled port of the viscous / turbulence operators exists) on a
    bounded sample: the same wing recipe at octree level 5."""
    import oracle
    from oracle import cfd as ocfd, euler as oeuler
    from immersedboundary_jl_b200 import synthetic
    OM = oracle.mesher
    cores = os.cpu_count() or 1
    pts, tri = synthetic.swept_wing(n_chord=32, n_span=32)
    h = F32(32.0 / 2 ** 5 / 8 * 1.01)
    lo, w = pts.min(axis=0) - 0.05, pts.max(axis=0) - pts.min(axis=0) + 0.1
    msh = OM.Mesh([-16, -16, -16], [32, 32, 32], ("wall", OM.Stereolitography(pts, tri), h), refinement_regions=[(OM.Box(lo.tolist(), w.tolist()), h)])
    fams = [("farfield", [(d, s) for d in range(3) for s in (False, True)])]
    dom = oracle.domain.Domain(msh, max_partition_size=max(4096, len(msh) // cores + 1), hypercube_families=fams)
    fl = ocfd.Fluid()
    a = np.sqrt(1.4 * 283.0 * 288.15)
    Pinf = np.array([101325.0, 288.15, 0.5 * a, 0.0, 0.0], F32)
    bcs = [("wall", ocfd.FlowBC(fl, Pinf[:3] * np.array([1, 1, 0], F32), normal_flow=True)), ("farfield", ocfd.FlowBC(fl, Pinf))]
    Q = synthetic.primitive2state_host(synthetic.euler_state(dom.centers))
    qR = (Q[:, 0] * F32(4.5e-5)).astype(F32)
    R, RR, cf = np.zeros_like(Q), np.zeros(len(Q), F32), np.zeros(len(Q), F32)
    res = oeuler.rans_residual(fl)
    n = len(Q)

    def step():
        oeuler.euler_ghost_update(dom, fl, Q, bcs)
        oeuler.rans_ghost_update(dom, Q, qR, [("wall", 0.0), ("farfield", 4.5e-5)])
        dom(res, Q, qR, R, RR, cf, n_threads=cores)

    step()
    t0 = time.perf_counter()
    steps = 3
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"NumPy restatement of the reference operators composed as oracle/euler.py: rans_residual (Julia cannot run here), the "
                      f"wing recipe at octree level 5 -> {n} cells in {len(dom.partitions)} partitions, {steps} evaluations, {cores} threads"}


if __name__ == "__main__":
    import ctypes as C
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"], help="c4 = the headline sphere Euler workload; c5 = wing RANS")
    ap.add_argument("--cells", type=int, default=0, help="override the target cell count (debugging)")
    ap.add_argument("--radius", type=float, default=0.0, help="override the refinement-ball radius (debugging)")
    ap.add_argument("--level", type=int, default=10, help="finest octree level (10 = the C4 workload; lower for smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the arithmetic = 1 pricing block")
    ap.add_argument("--analytic-sphere", action="store_true", help="exact sphere instead of the icosphere STL wall (quick smoke runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)
