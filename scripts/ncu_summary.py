"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with duration, DRAM read / write,
issue utilisation, registers, achieved occupancy.  Usage: python scripts/ncu_summary.py raw.csv > summary.txt"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "MB_rd", 1.0), ("dram__bytes_write.sum", "MB_wr", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("sm__inst_executed.sum", "Minst", 1e-6), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf", 1.0)]


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = {name: hdr.index(name) for name, _, _ in COLS if name in hdr}
    print("%-44s " % "kernel" + " ".join("%10s" % short for name, short, _ in COLS if name in idx))
    for r in rows[2:]:
        vals = []
        for name, short, scale in COLS:
            if name not in idx:
                continue
            v, u = r[idx[name]].replace(",", ""), units[idx[name]]
            try:
                f = float(v) * scale
                if u == "ns":
                    f /= 1e3
                if u == "byte":
                    f /= 1e6
                if u == "Kbyte":
                    f /= 1e3
                if u == "Gbyte":
                    f *= 1e3
                if u == "ms":
                    f *= 1e3
                vals.append("%10.2f" % f)
            except ValueError:
                vals.append("%10s" % v[:10])
        print("%-44s " % r[ki][:44] + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
