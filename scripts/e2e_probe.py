"""GPU probe: where the end-to-end step loses time against the device-resident step."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lrpx.pipeline import BatchExplainer
args = argparse.Namespace(images=64, words=19, vocab=10000, chunk=128)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
model, ex, imgs_h, toks_h = bench.build_problem(args, dev, 0)
imgs_h, toks_h = imgs_h.pin_memory(), toks_h.pin_memory()
imgs_d, toks_d = imgs_h.to(dev), toks_h.to(dev)
Q = 64 * 19
heat = torch.empty(Q, 3, 224, 224, device=dev)
heat_h = torch.empty(Q, 3, 224, 224).pin_memory(); words_h = torch.empty(Q, 19).pin_memory()
def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for graph in (True, False):
    pipe = BatchExplainer(ex, chunk=128, use_graph=graph)
    print(f"graph={graph}: device-resident {timeit(lambda: pipe.explain(imgs_d, toks_d, out=heat)):.2f} ms | "
          f"H2D only {timeit(lambda: pipe.explain(imgs_h, toks_h, out=heat)):.2f} | "
          f"D2H only {timeit(lambda: pipe.explain(imgs_d, toks_d, out=heat, host_out=(heat_h, words_h))):.2f} | "
          f"both {timeit(lambda: pipe.explain(imgs_h, toks_h, out=heat, host_out=(heat_h, words_h))):.2f}")
