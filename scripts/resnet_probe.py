"""GPU probe: encoder LRP of ResNet101 / VGG16 through the fp32 rule kernels (LRPtools.compute_lrp), batch of explanations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from LRPtools import lrp_wrapper
from models import resnet, vgg
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
for name, net, cshape in (("resnet101", resnet.resnet101(), (2048, 7, 7)), ("vgg16 features[0:-1]", vgg.vgg16(pretrained=False).features[0:-1], (512, 14, 14))):
    net = net.to(dev).eval()
    lrp_wrapper.add_lrp(net)
    for n in [int(v) for v in os.environ.get("BATCHES", "1,8").split(",")]:
        x = torch.randn(n, 3, 224, 224, device=dev); tgt = torch.randn(n, *cshape, device=dev)
        for _ in range(2): net.compute_lrp(x.clone(), target=tgt)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): net.compute_lrp(x.clone(), target=tgt)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"{name}: batch {n}: {dt * 1e3:.1f} ms per compute_lrp = {n / dt:.1f} explanations/s (fp32 rule kernels)")
