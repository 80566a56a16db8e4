#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: per kernel the launch shape, duration, DRAM traffic and throughput,
issue / occupancy figures, the shared-memory and L2 ATOMIC counters the north star asks for, and the stall mix with
the hottest SASS instructions (each counted once).   python tools/ncu_summary.py report.ncu-rep [kernel_regex ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("registers/thread", "launch__registers_per_thread"),
    ("dynamic smem/block", "launch__shared_mem_per_block_dynamic"), ("waves/SM", "launch__waves_per_multiprocessor"),
    ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("eligible warps / cycle", "smsp__warps_eligible.avg.per_cycle_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("shared ATOMIC warp instructions", "smsp__inst_executed_op_shared_atom.sum"),
    ("shared atomic wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum"),
    ("shared atomic bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum"),
    ("shared load wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
    ("shared store wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
    ("LSU data pipe % of peak", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("global ATOMIC / RED warp instructions", "smsp__inst_executed_op_global_atom.sum"),
    ("L2 atomic sectors", "lts__t_sectors_op_atom.sum"), ("L2 reduction sectors", "lts__t_sectors_op_red.sum"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("report:", rep.split("/")[-1])
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0]
        if short in seen:
            continue
        seen.add(short)
        print("\n== %s" % name[:150])
        dur = None
        for label, metric in WANT:
            if metric in hdr:
                i = hdr.index(metric)
                print("  %-40s %s %s" % (label, r[i], units[i]))
                if metric == "gpu__time_duration.sum":
                    dur = float(r[i].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[i], 1e-3)
        a = hdr.index("smsp__inst_executed_op_shared_atom.sum") if "smsp__inst_executed_op_shared_atom.sum" in hdr else None
        if a is not None and dur:
            print("  %-40s %.1f G lane-atomics/s upper bound (32 lanes)" % ("shared atomic rate", float(r[a].replace(",", "")) * 32 / dur / 1e9))
        sass = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_sass.py"), rep, short.split("::")[-1].split("<")[0], "8"],
                              capture_output=True, text=True).stdout
        print("  " + sass.replace("\n", "\n  ").rstrip())


if __name__ == "__main__":
    main()
