"""Summarise an .ncu-rep (ncu --set full) per launch: duration, DRAM bytes, pipe/issue utilisation, occupancy, top stalls.
usage: python tools/ncu_summary.py file.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}

def g(r, name, default=None):
    i = ci.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return r[i]

def scale(name, v):
    u = units[ci[name]] if name in ci else ""
    return v, u

stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_")]
out = []
for r in body:
    name = r[ci["Kernel Name"]][:40]
    dur = g(r, "gpu__time_duration.sum")
    du = units[ci["gpu__time_duration.sum"]]
    dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du, 1)
    def bytes_of(n):
        v = g(r, n, 0.0); u = units[ci[n]] if n in ci else "byte"
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    stalls = sorted(((g(r, c, 0.0) or 0.0, c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for c in stall_cols), reverse=True)[:4]
    rec = {
        "kernel": name, "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]], "us": round(dur_us, 1),
        "dram_MB": round((rd + wr) / 1e6, 1), "dram_GBps": round((rd + wr) / dur_us / 1e3, 1),
        "regs": g(r, "launch__registers_per_thread"),
        "occ_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_pct": g(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or g(r, "smsp__issue_active.avg.pct"),
        "ipc": g(r, "sm__inst_executed.avg.per_cycle_active"),
        "sm_thr_pct": g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "mem_thr_pct": g(r, "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit": g(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit": g(r, "lts__t_sector_hit_rate.pct"),
        "lsu_pct": g(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") or g(r, "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed"),
        "fma_pct": g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "alu_pct": g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "tensor_pct": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "stalls": ["%s=%.1f" % (n, v) for v, n in stalls],
    }
    out.append(rec)
    print(json.dumps(rec))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
