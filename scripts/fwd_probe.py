"""GPU probe: per-call CUDA-event time of the encoder forward (+gains) for a batch of images."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from lrpx import tc, _lib

n = int(os.environ.get("IMAGES", "64"))
sd = synth.vgg_state(1)
eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")], synth.VGG16_CFG, "cuda",
                     precision=os.environ.get("PRECISION", "bf16"), rule=os.environ.get("RULE", "alpha_beta"))
x = torch.randn(n, 3, 224, 224, device="cuda")
for _ in range(2):
    eng.forward(x)
torch.cuda.synchronize()
# wrap the C-ABI calls with events
recs = []
orig_check = _lib.check
import lrpx.tc as T
def timed(name):
    fn = getattr(_lib.lib(), name)
    def w(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = fn(*a); e1.record()
        recs.append((name, e0, e1))
        return rc
    return w
class L:
    def __getattr__(self, k):
        return timed(k)
T.lib = lambda: L()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.forward(x); e1.record()
torch.cuda.synchronize()
tot = 0.0
for name, a, b in recs:
    ms = a.elapsed_time(b); tot += ms
    print(f"{name:28s} {ms:8.3f} ms")
print(f"sum {tot:.3f} ms, forward wall (events) {e0.elapsed_time(e1):.3f} ms, ideal @1393.9 TF {eng.flops_forward_per_image() * n / 1393.9e12 * 1e3:.3f} ms")
