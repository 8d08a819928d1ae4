#!/usr/bin/env python
"""scripts/ncu_summary.py -- turn an .ncu-rep into the short text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of {rep} (per launch; cold-cache, serialised: compare shares, not absolutes)")
    for r in rows[2:]:
        print("\n== " + r[idx["Kernel Name"]][:150])
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
        for h in hdr:
            if "pipe_tensor" in h and h not in WANT and ("pct_of_peak_sustained_active" in h):
                print(f"  {h:72s} {r[idx[h]]} {units[idx[h]]}")
        stalls = []
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[idx[h]].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.3:
                    stalls.append((v, h.split("issue_stalled_")[1].split("_per_issue")[0]))
        print("  stall reasons (warps per issue, >= 0.3): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main()
