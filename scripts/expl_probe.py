"""GPU probe: one explainer forward (64 images x 19 steps), for an ncu launch list."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
args = argparse.Namespace(images=64, words=19, vocab=10000, chunk=128)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
model, ex, imgs, toks = bench.build_problem(args, dev, 0)
feat = torch.rand(64, 196, 512, device=dev)
toks = toks.to(dev)
for _ in range(3):
    ex.explainer_forward(feat, toks)
torch.cuda.synchronize()
