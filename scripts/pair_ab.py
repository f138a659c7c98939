"""GPU probe: A/B of the cta_group::2 pair mode per chain layer, alternating the two modes inside one process (the mode is
read from LRPX_TC_PAIR at every lrpx_tc_conv call), 128-request chunks (layers 0-4) / 1216 requests (others)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lrp-imagecaptioning-pytorch_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, synth
from lrpx import tc
dev = "cuda"
sd = synth.vgg_state(2000)
eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")], synth.VGG16_CFG, dev)
st = eng.forward(torch.randn(1, 3, 224, 224, device=dev))
L = len(eng.convs)
modes = sys.argv[1:] or ["0", "1"]
for li in range(L - 1, 0, -1):
    c, below = eng.convs[li], eng.convs[li - 1]
    n = 128 if li <= 4 else 1216
    rimg = torch.zeros(n, dtype=torch.int32, device=dev)
    a = torch.randn(tc.pf_rows(n, c.h, c.w), c.cout, device=dev).to(torch.bfloat16)
    oh, ow = (2 * c.h, 2 * c.w) if below.pool_after else (c.h, c.w)
    bufs = [None, torch.empty(tc.pf_rows(n, oh, ow) * c.cin, device=dev, dtype=torch.bfloat16)]
    res, outs = {}, {}
    for rep in range(3):
        for m in modes:
            os.environ["LRPX_TC_PAIR"] = m
            eng._run_layers(st, a, n, rimg, li, li + 1, bufs, 0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): eng._run_layers(st, a, n, rimg, li, li + 1, bufs, 0)
            e1.record(); torch.cuda.synchronize()
            res.setdefault(m, []).append(e0.elapsed_time(e1) / 10)
            if rep == 0: outs[m] = bufs[1].clone()
    same = all(torch.equal(outs[modes[0]], outs[m]) for m in modes)
    fl = 2.0 * n * c.h * c.w * c.cin * c.cout * 9
    print(f"layer {li:2d} hw {c.h:3d} {c.cout:3d}->{c.cin:3d} {'unpool' if below.pool_after else '      '} n={n:4d}: " +
          "  ".join(f"PAIR={m}: {min(res[m]):.4f} ms ({fl / min(res[m]) / 1e9:.0f} TF)" for m in modes) + f"  bit-identical: {same}")
