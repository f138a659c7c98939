"""GPU probe: ResNet101 encoder LRP throughput, batch 64 images x 19 requests each: bf16 tensor-core chain
(lrpx.tc_resnet) vs the fp32 CUDA-core rule kernels through LRPtools.compute_lrp."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lrp-imagecaptioning-pytorch_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from models import resnet
from lrpx import tc_resnet
dev = "cuda"
torch.manual_seed(0)
net = resnet.resnet101().to(dev).eval()
for m in net.modules():                      # non-trivial BatchNorm statistics
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
eng = tc_resnet.TcResNetEngine(net, dev)
B, T = int(os.environ.get("B", "64")), 19
x = torch.randn(B, 3, 224, 224, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
st = eng.forward(x); torch.cuda.synchronize()
e0, e1 = ev(), ev(); e0.record()
for _ in range(3): st = eng.forward(x)
e1.record(); torch.cuda.synchronize()
fwd = e0.elapsed_time(e1) / 3
feats = eng.features(st, "pixel")
Q = B * T
rimg = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(T)
r = torch.randn(Q, 49, 2048, device=dev) * 1e-3 * feats[rimg.long()]
heat = torch.empty(Q, 3, 224, 224, device=dev)
eng.relevance(st, r, rimg, out=heat); torch.cuda.synchronize()
e0, e1 = ev(), ev(); e0.record()
for _ in range(3): eng.relevance(st, r, rimg, out=heat)
e1.record(); torch.cuda.synchronize()
rel = e0.elapsed_time(e1) / 3
gf = eng.flops_per_explanation() / 1e9
print(f"ResNet101 bf16 chain: forward+gains {fwd:.1f} ms / {B} images; relevance {rel:.1f} ms / {Q} requests = "
      f"{Q / rel * 1e3:.0f} explanations/s ({gf:.1f} GFLOP each -> {gf * Q / rel:.0f} TFLOP/s); with the forward "
      f"{Q / (rel + fwd) * 1e3:.0f} explanations/s; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
assert torch.isfinite(heat).all()
