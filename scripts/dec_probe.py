"""GPU probe: one decoder-relevance call (1216 requests), for an ncu launch list."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lrpx import ops
args = argparse.Namespace(images=64, words=19, vocab=10000, chunk=128)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
model, ex, imgs, toks = bench.build_problem(args, dev, 0)
feat = torch.rand(64, 196, 512, device=dev)
toks = toks.to(dev)
st = ex.explainer_forward(feat, toks)
W = ex._lrp_weights()
B, T = 64, 19
req_img = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(T)
req_t = torch.arange(T, dtype=torch.int32, device=dev).repeat(B)
req_word = toks[:, 1:].reshape(-1).to(torch.int32)
torch.cuda.synchronize()
print("MARK")
for _ in range(3):
    ops.gridtd_decoder_lrp(st, W, req_img, req_t, req_word, tc_gemm=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gridtd_decoder_lrp(st, W, req_img, req_t, req_word, tc_gemm=True)
e1.record(); torch.cuda.synchronize()
print(f"gridtd_decoder_lrp, 1216 requests, eager launches: {e0.elapsed_time(e1) / 10:.3f} ms")
