"""GPU probe: the image-ablation and word-ablation experiments of evaluation.py:82-290 for every
(image, word) request of the bench workload (64 images x 19 words, V = 10000) in one batched pass."""
import os, sys, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lrpx.pipeline import BatchExplainer
from lrpx.ablation import AblationExperiments
B, T = int(os.environ.get("B", "64")), 19
args = argparse.Namespace(images=B, words=T, vocab=10000, chunk=128)
dev = torch.device("cuda")
model, ex, imgs, toks = bench.build_problem(args, dev, 0)
imgs, toks = imgs.to(dev), toks.to(dev)
be = BatchExplainer(ex, chunk=128)
ab = AblationExperiments(ex, chunk=128)
req_img = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(T)
req_t = torch.arange(T, dtype=torch.int32, device=dev).repeat(B)
eng = ex.engine()


def run():
    heat, r_words = be.explain(imgs, toks)
    feat = eng.features(eng.forward(imgs), "pixel").clone()
    pred = ex.explainer_forward(feat, toks)["pred"]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ab.image_ablation(imgs, toks, heat, req_img, req_t, pred)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sel = (req_t >= 6).nonzero()[:, 0]
    diff = ab.word_ablation(feat, toks, r_words[sel], req_img[sel], req_t[sel], pred)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return t1 - t0, t2 - t1, out, int(sel.numel())


run()
ti, tw, out, nw = run()
Q = B * T
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
heat = be.explain(imgs, toks)[0]
e0.record(); from lrpx import ops; ops.block_image(heat, 20, 8, images=imgs, req_img=req_img, want_mask=False); e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
gb = (heat.numel() * 4 * 2 + imgs.numel() * 4) / 1e9
print(f"image ablation, {Q} requests (mask, re-encode, beam search 3 x 20, teacher-forced scores): {ti * 1e3:.1f} ms = {Q / ti:.0f} ablations/s; "
      f"{int(out['disappear'].sum())} words disappeared")
print(f"word ablation, {nw} requests (t >= 6): {tw * 1e3:.1f} ms")
print(f"lrpx_block_image_f32: {ms:.3f} ms for {Q} requests = {gb / ms * 1e3:.0f} GB/s (read heat + image, write masked image)")
