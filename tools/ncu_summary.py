#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the per-kernel summaries kept under profiles/:
   python tools/ncu_summary.py gpurun_out/r1_k1t.ncu-rep frontend_tma_kernel profiles/r1_ncu_frontend_tma.csv
writes one `metric,unit,value` row per raw-page metric of the FIRST launch whose name matches."""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "dram__", "gpu__", "l1tex__data_bank", "l1tex__t_bytes_pipe_lsu_mem_global_op_ldgsts",
        "launch__", "sm__cycles_elapsed", "sm__inst_executed_pipe", "sm__throughput", "sm__warps_active", "smsp__average",
        "smsp__cycles_active", "smsp__inst_executed.sum", "smsp__issue_active", "lts__t_bytes", "l1tex__t_sector_hit_rate",
        "smsp__pcsamp", "sm__sass_inst_executed_op_shared", "smsp__inst_executed_op_shared")


def main():
    rep, kern, out = sys.argv[1:4]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if kern in r[ki])
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, row):
            if any(h.startswith(k) for k in KEEP):
                w.writerow([h, u, v])
    print("wrote", out)


if __name__ == "__main__":
    main()
