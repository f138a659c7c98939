"""GPU probe: lrpx_tc_maxpool2_bf16 on the four VGG16 pools of 64 images (warm L2 flushed by the 1 GB working set)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200"))
import torch
from lrpx import tc
n = 64
for h, c in ((224, 64), (112, 128), (56, 256), (28, 512)):
    rows = tc.pf_rows(n, h, h)
    act = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
    gain = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
    for _ in range(3): tc.maxpool2(act, gain, n, h, h, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): tc.maxpool2(act, gain, n, h, h, c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = rows * c * 2 * 2 + tc.pf_rows(n, h // 2, h // 2) * c * 5
    print(f"maxpool2 {h}->{h // 2}, {c} ch: {ms:.4f} ms, {nbytes / 1e6:.0f} MB -> {nbytes / 1e9 / (ms * 1e-3):.0f} GB/s")
