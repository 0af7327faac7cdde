#!/usr/bin/env python
"""One train step out of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list of `bench.py --no-graph`: per-kernel and per-bench-kind time and DRAM traffic.

    python scripts/ncu_step_summary.py gpurun_out/launches.csv profiles/rNN_launches_summary.txt profiles/ncu_traffic.json

The step is the span between two consecutive stem_fwd launches.  ncu times are cold-cache and serialised:
compare SHARES with bench.py's live per-kind numbers, not absolute values.  The JSON maps the bench kinds
(plan.cu kProfNames) to DRAM bytes per launch; bench.py reports it as `roofline.traffic` for the dominant kind.
"""
import collections
import csv
import json
import re
import sys

# kernel-name pattern -> bench kind (first match wins).  A dw_bwd "launch" of bench.py is one layer = the
# data-gradient kernel + the weight-gradient kernel, so bytes are summed over both and divided by the layers.
KINDS = [
    (r"stem_fwd", "stem_fwd"), (r"stem_wgrad", "stem_wgrad"),
    (r"gemm_tn_", "gemm_wgrad"), (r"gemm_nt_", "gemm_nt"),
    (r"dwc_fwd|d2_fwd|ww_conv_kernel", "dw_fwd"),
    (r"dwc_bwd", "dw_bwd"),
    (r"bn_.*finalize", "bn_finalize"), (r"apply_xform", "apply_xform"), (r"affine2", "affine2"),
    (r"act_bwd_stats", "act_bwd_stats"), (r"se_", "se_fc"), (r"heads_", "heads"), (r"pool_finalize", "pool"),
    (r"optim_kernel", "optimizer"), (r"pack_table", "pack_weights"),
]
CONV = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines[hi:]):
        d = by.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        elif r["Metric Name"] == "dram__bytes_read.sum":
            d["rd"] = v * CONV.get(u, 1.0)
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["wr"] = v * CONV.get(u, 1.0)
    return list(by.values())


def kind_of(name):
    for pat, k in KINDS:
        if re.search(pat, name):
            return k
    return "other"


def main(path, out_txt, out_json):
    ls = load(path)
    starts = [i for i, d in enumerate(ls) if "stem_fwd" in d["name"]]
    if len(starts) >= 2:
        step = ls[starts[0]:starts[1]]
        note = "one full step (between two stem_fwd launches)"
    else:
        step = ls[starts[0]:] if starts else ls
        note = "PARTIAL step (capture window ended early)"
    per_kernel = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    per_kind = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in step:
        nm = re.sub(r"^void ", "", re.sub(r"\(td3d::.*|\(CUtensorMap.*|\(const .*|\(float.*|\(.*\)$", "", d["name"]))
        nm = nm.replace("td3d::", "").replace("<unnamed>::", "")
        for agg, key in ((per_kernel, nm), (per_kind, kind_of(d["name"]))):
            a = agg[key]
            a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("rd", 0.0); a[3] += d.get("wr", 0.0)
    tot = sum(a[1] for a in per_kernel.values())
    with open(out_txt, "w") as f:
        f.write(f"# {path}: {len(step)} launches, {tot / 1e3:.2f} ms summed kernel time, {note}\n")
        f.write("# ncu per-launch times are cold-cache and serialised: compare shares, not absolutes\n")
        f.write(f"{'kernel':64s} {'n':>4s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s}\n")
        for k, a in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:64]:64s} {a[0]:4d} {a[1]:10.1f} {a[1] / tot:6.3f} {a[1] / a[0]:8.1f} {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f}\n")
        f.write("\n# by bench kind\n")
        for k, a in sorted(per_kind.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:16s} n={a[0]:4d} {a[1]:10.1f} us  share {a[1] / tot:5.3f}  dram {(a[2] + a[3]) / 1e6:9.1f} MB\n")
    layers = {"dw_bwd": 15, "dw_fwd": 15}      # bench counts one launch per depthwise layer
    traffic = {}
    for k, a in per_kind.items():
        n = layers.get(k, a[0]) if k in layers and a[0] >= layers[k] else a[0]
        traffic[k] = {"dram_bytes_per_launch": (a[2] + a[3]) / max(n, 1), "launches": n, "kernel_launches": a[0],
                      "dram_bytes_per_step": a[2] + a[3], "ncu_us_per_step": a[1]}
    # bench.py splits gemm_nt into fwd / dgrad; both map to the same kernel
    if "gemm_nt" in traffic:
        traffic["gemm_fwd"] = traffic["gemm_nt"]
        traffic["gemm_dgrad"] = traffic["gemm_nt"]
    traffic["_source"] = f"{path} ({note}); dram__bytes_read.sum + dram__bytes_write.sum per launch"
    with open(out_json, "w") as f:
        json.dump(traffic, f, indent=1)
    print(open(out_txt).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
