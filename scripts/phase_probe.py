"""GPU probe: time of each phase of the batched explanation step when replayed from its own CUDA graph."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lrpx import ops

args = argparse.Namespace(images=int(os.environ.get("IMAGES", 64)), words=19, vocab=10000, chunk=128)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
model, ex, imgs, toks = bench.build_problem(args, dev, 0)
imgs, toks = imgs.to(dev), toks.to(dev)
eng, W = ex.engine(), ex._lrp_weights()
B, T = args.images, args.words
req_img = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(T)
req_t = torch.arange(T, dtype=torch.int32, device=dev).repeat(B)
req_word = toks[:, 1:].reshape(-1).to(torch.int32)
heat = torch.empty(B * T, 3, 224, 224, device=dev)
state = {}
def p_fwd():
    state["est"] = eng.forward(imgs); state["feat"] = eng.features(state["est"], "pixel")
def p_expl():
    state["st"] = ex.explainer_forward(state["feat"], toks)
def p_dec():
    state["r_feat"], state["r_words"] = ops.gridtd_decoder_lrp(state["st"], W, req_img, req_t, req_word, tc_gemm=True)
def p_chain():
    eng.relevance(state["est"], state["r_feat"], req_img, chunk=args.chunk, out=heat)
phases = [("encoder_forward_gains", p_fwd), ("explainer_forward", p_expl), ("decoder_relevance", p_dec), ("encoder_relevance_chain", p_chain)]
for _, f in phases:
    f()
torch.cuda.synchronize()
tot = 0
for name, f in phases:
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5; tot += ms
    print(f"{name:28s} graph replay {ms:8.3f} ms")
print(f"sum {tot:.3f} ms")
