#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python scripts/ncu_summary.py list  gpurun_out/launches.csv profiles/rNN_launches.md  "<command line>"
  python scripts/ncu_summary.py full  gpurun_out/prof.ncu-rep profiles/rNN_tc_conv.md   "<command line>"
"""
import collections
import csv
import re
import subprocess
import sys


def launch_list(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot, n = 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v; tot += v; n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n"
                f"{n} launches, {tot:.3f} ms total device time (cold-cache, serialised: compare SHARES).\n\n"
                "| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        ours = 0.0
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if t / tot < 0.002:
                continue
            f.write(f"| `{k}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |\n")
        for k, (c, t) in agg.items():
            if "lrpx::" in k:
                ours += t
        f.write(f"\nlrpx kernels: {100 * ours / tot:.1f}% of device time; "
                f"tc_conv_*kernel (all epilogues, forward + chain + decoder GEMMs): {100 * sum(t for k, (c, t) in agg.items() if 'tc_conv' in k) / tot:.1f}%\n")
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size"]


def full(src, dst, cmd):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture (tc_conv_kernel)\n\ncommand: `{cmd}`\n\n| # | kernel | " +
                " | ".join(f"{c} [{units[idx[c]]}]" for c in cols) + " |\n|" + "---|" * (len(cols) + 2) + "\n")
        tot_r = tot_w = tot_t = 0.0
        # explanations covered by each captured launch, e.g. "1216x8,128x5" (8 launches over 1216 requests, then 5 over 128)
        per_launch = []
        if len(sys.argv) > 5:
            for part in sys.argv[5].split(","):
                n, _, c = part.partition("x")
                per_launch += [int(n)] * int(c or 1)
        bytes_per_expl = 0.0
        for i, r in enumerate(rows[2:]):
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
            f.write(f"| {i} | `{name}` | " + " | ".join(r[idx[c]] for c in cols) + " |\n")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot_r += float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
            tot_w += float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
            tot_t += float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[idx["gpu__time_duration.sum"]]]
            if i < len(per_launch):
                bytes_per_expl += (float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]] +
                                   float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]) / per_launch[i]
        f.write(f"\nall {len(rows) - 2} launches: dram read {tot_r / 1e6:.1f} MB + write {tot_w / 1e6:.1f} MB = "
                f"{(tot_r + tot_w) / 1e6:.1f} MB in {tot_t:.1f} us\n")
    # machine-readable copy for bench.py's roofline.traffic
    import json
    json.dump({"launches": len(rows) - 2, "dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "gpu_time_us": tot_t,
               "explanations_per_launch": per_launch, "dram_bytes_per_explanation": bytes_per_expl, "command": cmd,
               "source": dst}, open(dst.replace(".md", ".json"), "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    {"list": launch_list, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
