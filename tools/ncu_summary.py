"""Condense an ``ncu --set full`` report into the per-kernel table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_ncu_full_summary.csv [kernel-regex]

Reads the report with ``ncu -i ... --page raw --csv`` (works without a GPU) and keeps the columns the
roofline discussion in DESIGN.md uses.  The last launch of every kernel name is kept (the first one is
the warm-up pass of tools/profile_case.py).
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(k) for k in KEEP if k in head]
    last = {}
    for r in body:
        name = r[head.index("Kernel Name")]
        if pat is None or pat.search(name):
            last[name] = r
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in last.values():
            w.writerow([r[i] for i in idx])
    print(f"{len(last)} kernels -> {out}")


if __name__ == "__main__":
    main()
