"""GPU probe: accuracy of the fp32-accurate chain (precision='fp32') against the reference fixtures and the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lrp-imagecaptioning-pytorch_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import synth, lrp_oracle as O
from lrpx import tc
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
def gold(name):
    d = np.load(os.path.join(ROOT, "tests/golden", name + ".npz")); return {k: torch.from_numpy(d[k]) if d[k].ndim else d[k] for k in d.files}
def report(tag, a, b):
    a, b = a.cpu().double(), b.cpu().double()
    err = (a - b).abs(); sc = float(b.abs().max())
    print(f"{tag}: max err/max {float(err.max())/sc:.3e}  rel L2 {float((a-b).norm()/b.norm()):.3e}  "
          f"frac > 1e-4*(|b|+max) {float((err > 1e-4*(b.abs()+sc)).double().mean()):.3e}  frac > 1e-4|b|+1e-6*max {float((err > 1e-4*b.abs()+1e-6*sc).double().mean()):.3e}")
for name, size in (("vgg16_64", 64), ("vgg16_224", 224)):
    g = gold(name); seed = int(g["seed"]); sd = synth.vgg_state(seed)
    ws = [sd[k] for k in sd if k.endswith("weight")]; bs = [sd[k] for k in sd if k.endswith("bias")]
    if size == 64:
        x, tgt = g["x"], g["target"]
    else:
        gen = torch.Generator().manual_seed(seed + 1000)
        x = torch.randn(1, 3, 224, 224, generator=gen); tgt = torch.randn(1, 512, 14, 14, generator=gen) * 1e-3
    layers64 = [tuple(v.double() if torch.is_tensor(v) else v for v in l) for l in O.vgg_layers_from_state(sd)]
    acts64 = O.sequential_forward(layers64, x.double())
    ref64 = O.sequential_lrp(layers64, x.double(), tgt.double())
    for prec in ("fp32", "bf16"):
        eng = tc.TcVggEngine(ws, bs, synth.VGG16_CFG, DEV, precision=prec)
        st = eng.forward(x.to(DEV))
        report(f"{name} [{prec}] features vs fp64", eng.features(st, "nchw"), acts64[-1])
        heat = eng.relevance(st, tgt.flatten(2).transpose(1, 2).contiguous().to(DEV))
        report(f"{name} [{prec}] heat vs fixture(fp32 ref)", heat, g["rel"])
        report(f"{name} [{prec}] heat vs fp64 oracle", heat, ref64)
    report(f"{name} fixture(fp32 ref) vs fp64 oracle", g["rel"], ref64)
    if "feats" in g: report(f"{name} fixture feats vs fp64", g["feats"], acts64[-1])
