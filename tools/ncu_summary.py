"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/ (read here, no GPU needed).

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max.per_second", "sm clock"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM rate"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "instructions"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}\n# columns: ncu --set full --clock-control none, one block per captured launch\n")
    for r in rows[2:]:
        print(f"kernel: {r[col['Kernel Name']]}")
        for key, label in KEYS:
            if key in col:
                print(f"  {label:24s} {r[col[key]]} {units[col[key]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
