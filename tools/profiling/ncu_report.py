"""Summarise .ncu-rep captures (tools/profiling/ncu_kernels.sh) as a markdown table: duration, tensor-pipe and DRAM
utilisation, achieved occupancy, registers / smem, top stall reasons.  Needs `ncu` on PATH (no GPU).

  python tools/profiling/ncu_report.py gpurun_out/*.ncu-rep > profiles/ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time us"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read MB"),
    ("dram__bytes_write.sum", "DRAM write MB"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header = rows[0]
    units = rows[1] if len(rows) > 1 and not rows[1][0].strip().isdigit() else [""] * len(header)
    data = [r for r in rows[1:] if len(r) == len(header) and r[0].strip().isdigit()]
    return header, units, data


TIME_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
BYTES_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}


def normalise(value, unit):
    """ncu scales every column's unit to its values: bring times to us and byte counts to MB."""
    try:
        f = float(value.replace(",", ""))
    except ValueError:
        return value
    if unit in TIME_TO_US:
        f *= TIME_TO_US[unit]
    elif unit in BYTES_TO_MB:
        f *= BYTES_TO_MB[unit]
    return str(int(f)) if f == int(f) else f"{f:.1f}"


def main(paths):
    print("| capture | kernel | " + " | ".join(n for _, n in METRICS) + " | top stalls (warps per issue) |")
    print("|---|---|" + "---|" * (len(METRICS) + 1))
    for path in paths:
        try:
            header, units, data = raw_rows(path)
        except Exception as e:  # noqa: BLE001 - report and continue with the other captures
            print(f"| {path} | (unreadable: {e}) |")
            continue
        col = {h: i for i, h in enumerate(header)}
        for r in data:
            kernel = r[col.get("Kernel Name", 4)][:48]
            cells = []
            for metric, _ in METRICS:
                i = col.get(metric)
                cells.append(normalise(r[i], units[i]) if i is not None else "-")
            stalls = []
            for h, i in col.items():
                if h.startswith(STALL_PREFIX) and h.endswith("_per_warp_active.pct") is False and "ratio" in h:
                    try:
                        stalls.append((float(r[i].replace(",", "")), h[len(STALL_PREFIX):].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            top = ", ".join(f"{n} {v:.2f}" for v, n in stalls[:4])
            print(f"| {path.split('/')[-1]} | {kernel} | " + " | ".join(cells) + f" | {top} |")


if __name__ == "__main__":
    main(sys.argv[1:])
