#!/usr/bin/env python
"""Compact markdown table of an `ncu --set full` report (read here on the CPU box: `ncu -i rep --page raw --csv`).
    python tools/ncu_table.py gpurun_out/x.ncu-rep > profiles/r02_x_ncu.md"""
import csv, io, subprocess, sys

COLS = [("gpu__time_duration.sum", "us", 1.0),
        ("sm__cycles_elapsed.avg.per_second", "SM GHz", 1.0),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
        ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0),
        ("launch__grid_size", "grid", 1.0),
        ("launch__block_size", "block", 1.0),
        ("smsp__inst_executed.sum", "warp inst (M)", 1e-6),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1.0)]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n, _ in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        name = name.replace("void ", "").replace("ac::", "").replace("(int)", "").replace("(bool)", "")
        name = name.split("(CUtensorMap")[0].split("(const ")[0][:70]
        cells = []
        for key, _, scale in COLS:
            if key not in idx:
                cells.append("-")
                continue
            v = r[idx[key]].replace(",", "")
            u = units[idx[key]]
            try:
                f = float(v) * scale
                if key.startswith("dram__bytes"):
                    f = f if u == "Mbyte" else (f / 1e6 if u == "byte" else (f * 1e3 if u == "Gbyte" else (f / 1e3 if u == "Kbyte" else f)))
                if key == "gpu__time_duration.sum":
                    f = f if u == "us" else (f * 1e3 if u == "ms" else (f / 1e3 if u == "ns" else f))
                cells.append(f"{f:.1f}" if abs(f) < 1e5 else f"{f:.3g}")
            except ValueError:
                cells.append(v[:10])
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
