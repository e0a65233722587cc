"""Condense an Nsight Compute report into the handful of metrics DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/foo.ncu-rep > profiles/foo.md
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# {path}\n")
    print(f"{len(data)} captured launch(es); `ncu --set full --clock-control none`.\n")
    print("| metric | unit | " + " | ".join(f"#{i} {r[kn].split('(')[0][-38:]}" for i, r in enumerate(data)) + " |")
    print("|---|---|" + "---|" * len(data))
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
