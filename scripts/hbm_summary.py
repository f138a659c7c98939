#!/usr/bin/env python
"""profiles/rNN_hbm_kernels.md from the ncu launch list (gpurun_out/launches.csv): achieved GB/s of the HBM-bound kernels
of the bench step = algorithmic bytes (one read of each input, one write of each output; formulas below) / the launch's
gpu__time_duration (cold cache, serialised).  Workload: 64 images x 19 words, VGG16 224x224, H = 512, P = 196.
  python scripts/hbm_summary.py gpurun_out/launches.csv profiles/r1_hbm_kernels.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, T, P, H, C = 64, 19, 196, 512, 512
Q = B * T
pf = lambda n, h: n * (h + 1) * (h + 1)               # padded-flat rows of n images of h x h pixels


def pool_bytes(h, ch):                                  # maxpool2_pf: act + gain in (bf16), pooled act + gain (bf16) + 1-byte argmax out
    return pf(B, h) * ch * 2 * 2 + pf(B, h // 2) * ch * (2 + 2 + 1)


KERNELS = [
    ("maxpool2_pf_kernel", [("pool 224->112, 64 ch", pool_bytes(224, 64)), ("pool 112->56, 128 ch", pool_bytes(112, 128)),
                            ("pool 56->28, 256 ch", pool_bytes(56, 256)), ("pool 28->14, 512 ch", pool_bytes(28, 512))]),
    ("scale_rows_kernel", [("r_feat fp32 (Q,196,512) -> s bf16 PF", Q * P * C * 4 + pf(Q, 14) * C * 2)]),
    ("im2col3_split_kernel", [("x fp32 (64,3,224,224) -> 64-column bf16 PF rows", B * 3 * 224 * 224 * 4 + pf(B, 224) * 64 * 2)]),
    ("grid_attn_rows_kernel", [("G = A/stab(A_pre) fp32 (B,196,512) + uctx fp32 (Q,T,512) in, split bf16 operand (Q*196, 2*512 [hi|lo]) out",
                                B * P * H * 4 + Q * T * H * 4 + Q * P * 2 * H * 2)]),
    ("grid_attn_gain_kernel", [("A, A_pre fp32 in, G fp32 out (B,196,512)", 3 * B * P * H * 4)]),
]


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    dur = lambda r: float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    with open(dst, "w") as f:
        f.write("# HBM-bound kernels of the bench step: achieved GB/s from the ncu launch list\n\n"
                f"source: `{src}` (gpu__time_duration.sum, --clock-control none, cold cache); peak = measured copy bandwidth "
                f"{pk:.0f} GB/s (MEASURED_PEAKS.json). Algorithmic bytes = one read of each input + one write of each output.\n\n"
                "| kernel | launch | algorithmic MB | ms | GB/s | of peak |\n|---|---|---:|---:|---:|---:|\n")
        for name, cases in KERNELS:
            ds = [dur(r) for r in rows if name in r["Kernel Name"]]
            for i, (what, nbytes) in enumerate(cases):
                if i >= len(ds):
                    continue
                gbs = nbytes / 1e9 / (ds[i] * 1e-3)
                f.write(f"| `{name}` | {what} | {nbytes / 1e6:.1f} | {ds[i]:.4f} | {gbs:.0f} | {gbs / pk:.2f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
