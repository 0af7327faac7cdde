#!/usr/bin/env python
"""Table of an `ncu --set full` report (one line per captured launch): duration, DRAM traffic, issue utilisation,
occupancy, registers and the top warp-stall reasons.   python scripts/ncu_full_table.py report.ncu-rep > table.txt"""
import csv
import re
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def g(d, k):
        try:
            return float(d[ix[k]].replace(",", ""))
        except (ValueError, KeyError, IndexError):
            return float("nan")

    def mb(d, k):
        u = units[ix[k]]
        return g(d, k) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)

    stall_keys = [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    print(f"# {rep}: {len(data)} launches (ncu --set full --clock-control none; cold-cache, serialised)")
    print(f"{'kernel':58s} {'grid':>13s} {'us':>8s} {'rd_MB':>8s} {'wr_MB':>7s} {'GB/s':>7s} {'issue%':>6s} {'warps%':>6s} {'regs':>4s}  top stalls (warps per issue)")
    for d in data:
        nm = re.sub(r"\(td3d::.*|\(CUtensorMap.*|\(const .*|\(float.*", "", d[ix["Kernel Name"]])
        nm = nm.replace("void ", "").replace("td3d::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")[:58]
        us = g(d, "gpu__time_duration.sum")
        if units[ix["gpu__time_duration.sum"]] == "ns":
            us /= 1e3
        elif units[ix["gpu__time_duration.sum"]] == "ms":
            us *= 1e3
        rd, wr = mb(d, "dram__bytes_read.sum"), mb(d, "dram__bytes_write.sum")
        st = sorted(((g(d, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys), reverse=True)[:4]
        print(f"{nm:58s} {d[ix['Grid Size']]:>13s} {us:8.1f} {rd:8.1f} {wr:7.1f} {(rd + wr) / us * 1e3:7.0f} "
              f"{g(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{int(g(d, 'launch__registers_per_thread')):4d}  " + ", ".join(f"{n} {v:.2f}" for v, n in st))


if __name__ == "__main__":
    main(sys.argv[1])
