#!/usr/bin/env python
"""Summarise an ncu report (`ncu --set full ... -o X`) per launch: the metrics DESIGN.md and the
bench line refer to (duration, DRAM bytes, sectors per request, occupancy, stall reasons).

    python tools/ncu_summary.py gpurun_out/X.ncu-rep [title...] > profiles/Y.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "lg_throttle", "mio_throttle", "math_pipe_throttle",
          "membar", "not_selected", "branch_resolving", "no_instruction", "drain", "sleeping"]


def main():
    rep = sys.argv[1]
    title = " ".join(sys.argv[2:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s" % (title or rep))
    print("# source: %s (ncu --set full --clock-control none; per-launch values, cold caches, serialised)" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print("\nkernel: %s" % name[:150])
        vals = {}
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                vals[key] = (r[col[key]], units[col[key]])
                print("  %-48s %s %s" % (label, r[col[key]], units[col[key]]))

        def num(key):
            return float(vals[key][0].replace(",", "")) if key in vals else None
        rd, wr, dur = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
        if rd is not None and wr is not None and dur:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            b = rd * scale.get(vals["dram__bytes_read.sum"][1], 1.0) + wr * scale.get(vals["dram__bytes_write.sum"][1], 1.0)
            t = dur * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "s": 1.0,
                       "second": 1.0}.get(vals["gpu__time_duration.sum"][1], 1e-3)
            print("  %-48s %.3f GB  →  %.0f GB/s" % ("DRAM read + write", b / 1e9, b / t / 1e9))
        lr, ls = num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        if lr and ls:
            print("  %-48s %.2f" % ("sectors per global load request", ls / lr))
        stalls = []
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if key in col and r[col[key]] not in ("", "0"):
                stalls.append((float(r[col[key]]), s))
        if stalls:
            print("  warp stall reasons (warps per issue): " + ", ".join("%s %.2f" % (s, v) for v, s in sorted(stalls, reverse=True)[:6]))


if __name__ == "__main__":
    main()
