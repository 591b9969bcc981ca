"""Turn ncu artefacts from gpurun_out/ into small text summaries under profiles/ (tracked)."""
import collections, csv, subprocess, sys, os

def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        fh.write(f"# source: {path}\n")
        for k, v in agg.items():
            fh.write(f"{k:40s} launches={len(v):3d} avg_ms={sum(v)/len(v)/1e6:9.3f} share={sum(v)/tot*100:5.1f}%\n")

def launches_dram(path, out_txt, out_json, cells, note=""):
    """Launch list captured with gpu__time_duration.sum + dram__bytes_{read,write}.sum: per-kernel averages, and the
    DRAM bytes of one residual call (all kernels but the ghost update) as JSON for bench.py's roofline.traffic."""
    import json
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]; ki = H.index("Kernel Name"); mi = H.index("Metric Name"); vi = H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg.setdefault(k, collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(d["gpu__time_duration.sum"]) for d in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 python bench.py --steps 2 --warmup 1",
             "# (cold-cache, serialised launches: compare SHARES with the live CUDA-event timing of bench.py, not absolutes)"]
    if note: lines.append("# " + note)
    lines.append(f"{'kernel':42s} {'n':>3s} {'avg ms':>8s} {'share':>6s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
    rd_t = wr_t = t_t = 0.0
    for k, d in agg.items():
        t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        lines.append(f"{k:42s} {len(t):3d} {sum(t) / len(t) / 1e6:8.3f} {sum(t) / tot * 100:5.1f}% {sum(rd) / len(rd) / 1e6:11.1f} {sum(wr) / len(wr) / 1e6:11.1f}")
        if "ghost" not in k:
            rd_t += sum(rd) / len(rd); wr_t += sum(wr) / len(wr); t_t += sum(t) / len(t)
    lines.append(f"# one ibx_residual_euler call (everything but the ghost kernels): {t_t / 1e6:.3f} ms serialised, DRAM {rd_t / 1e9:.3f} GB read + "
                 f"{wr_t / 1e9:.3f} GB written = {(rd_t + wr_t) / cells:.1f} B per cell (algorithmic: 44 B per cell)")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump({"workload": "C4", "cells": cells, "kernel": "ibx_residual_euler (k_prim + sensor kernels + k_gen_faces + k_march_flux)",
               "dram_bytes_per_launch": rd_t + wr_t, "source": out_txt + " (ncu dram__bytes_read.sum + dram__bytes_write.sum, summed over the kernels of one call)"},
              open(out_json, "w"), indent=1)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads"]

def report(rep, out, cells=None, idx=0):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U, V = rows[0], rows[1], rows[2 + idx]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]   # one table per captured launch
    want = V[H.index('Kernel Name')].split('(')[0]
    pick = [k for k, i in enumerate(starts) if srows[i][1].replace('(int)', '').replace('(bool)', '').split('(')[0] == want] or [idx]
    k = pick[0]
    srows = srows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(srows))]
    SH = srows[1]
    si, ei = SH.index("Source"), SH.index("Instructions Executed")
    cols = [(i, h) for i, h in enumerate(SH) if h.startswith("stall_") and "Not Issued" not in h]
    op, st, tot = collections.Counter(), collections.Counter(), 0
    for r in srows[2:]:
        if r and r[0] in ("Kernel Name", "Address"): break   # multi-launch report: only the first launch is summarised
        if len(r) <= ei: continue
        m = r[si].strip().split()
        if not m: continue
        mn = (m[0] if not m[0].startswith("@") else m[1]).split(".")[0]
        n = int(r[ei]); op[mn] += n; tot += n
        for i, h in cols: st[h] += int(r[i])
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on; source: {rep}\n")
        fh.write(f"kernel: {V[H.index('Kernel Name')]}\n")
        for w in WANT:
            if w in H:
                i = H.index(w); fh.write(f"{w:72s} {U[i]:>16s} {V[i]}\n")
        fh.write(f"warp instructions executed (source page): {tot}\n")
        if cells:
            fh.write(f"thread-instructions per cell: {tot * 32 / cells:.1f}  (cells in this launch: {cells})\n")
            i = H.index("dram__bytes_read.sum"); j = H.index("dram__bytes_write.sum")
            def val(k):
                x = float(V[k].replace(",", "")); u = U[k]
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            fh.write(f"DRAM traffic per cell: {(val(i) + val(j)) / cells:.1f} B\n")
        fh.write("opcode mix (share of executed warp instructions):\n")
        for k, v in op.most_common(16): fh.write(f"  {k:8s} {v / tot * 100:5.1f}%\n")
        S = sum(st.values()) or 1
        fh.write("warp stall samples: " + ", ".join(f"{k[6:]} {v / S * 100:.1f}%" for k, v in st.most_common(9)) + "\n")

if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches": launches(sys.argv[2], sys.argv[3])
    elif cmd == "launches_dram": launches_dram(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), " ".join(sys.argv[6:]))
    else: report(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] != '-' else None, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
