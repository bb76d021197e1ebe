#!/usr/bin/env python
"""Text summary of an ncu --set full report, one block per captured kernel launch (profiles/r2_*_ncu_summary.txt).

    python pingpongpro_b200/tools/ncu_summary.py REPORT.ncu-rep > profiles/NAME.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_static", "static shared memory / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks/SM"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none, {len(rows) - 2} kernel launch(es)\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        print(f"## {name}   (launch id {r[col['ID']]})")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"  {label:48s} {r[col[key]]} {units[col[key]]}")
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_write.sum"]]]
            t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[units[col["gpu__time_duration.sum"]]]
            print(f"  {'DRAM traffic / duration':48s} {(rd + wr) / t / 1e9:.1f} GB/s  ({(rd + wr) / 1e6:.1f} MB)")
        except Exception:
            pass
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("  stalls (warps per issue, share): " + ", ".join(f"{n} {v:.2f} ({100 * v / tot:.0f} %)" for v, n in stalls[:6]))
        print()


if __name__ == "__main__":
    main()
