"""Probe: where does the tcgen05 'gradient' rule differ from the oracle's backward?  (test infrastructure)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", "lrp-imagecaptioning-pytorch_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import lrp_oracle as O
import synth
from lrpx import tc

DEV = "cuda"
for cfg in ([64], [64, 64], [64, 64, "M", 128], [64, 64, "M", 128, 128, "M", 256, "M", 256]):
    sd = synth.vgg_state(91, cfg)
    layers = O.vgg_layers_from_state(sd, cfg)
    x = synth.images(92, 2, 64)
    feat = O.sequential_forward(layers, x)[-1]
    rows = [0, 1, 1, 0, 1]
    tgt = torch.randn(len(rows), *feat.shape[1:], generator=torch.Generator().manual_seed(93))
    for guided in (False, True):
        want = torch.cat([O.sequential_gradient(layers, x[b:b + 1], tgt[q:q + 1], guided=guided) for q, b in enumerate(rows)])
        for prec in ("fp32", "bf16"):
            ws = [sd[k] for k in sd if k.endswith("weight")]
            bs = [sd[k] for k in sd if k.endswith("bias")]
            eng = tc.TcVggEngine(ws, bs, cfg, DEV, precision=prec, rule="guided" if guided else "gradient")
            st = eng.forward(x.to(DEV))
            f = eng.features(st, "nchw").cpu()
            got = eng.relevance(st, tgt.flatten(2).transpose(1, 2).contiguous().to(DEV),
                                torch.tensor(rows, dtype=torch.int32, device=DEV)).cpu()
            scale = want.abs().max()
            e = (got - want).abs() / scale
            print(f"cfg={cfg} guided={guided} {prec}: feat err {float((f - feat).abs().max() / feat.abs().max()):.2e} "
                  f"max {float(e.max()):.2e} relL2 {float((got - want).norm() / want.norm()):.2e} "
                  f"frac>1e-3 {float((e > 1e-3).float().mean()):.2e} frac>1e-4 {float((e > 1e-4).float().mean()):.2e}")
