"""GPU probe: latency of the reference-shaped per-image API (ExplainGridTDAttention.explain_caption, 1 image x 19 words)."""
import os, sys, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
args = argparse.Namespace(images=1, words=19, vocab=10000, chunk=128)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
model, ex, imgs, toks = bench.build_problem(args, dev, 0)
img = imgs[:1].to(dev)
tk = toks[0].tolist()
ex.preprocess_img = lambda p: img
model.beam_search = lambda *a, **k: (["caption"], tk[1:])        # teacher-force the synthetic caption (beam search is torch host code)
ex.visualize_explanations = lambda *a, **k: None
os.makedirs(ex.args.save_path, exist_ok=True)
for name, fn in (("explain_caption (19 words, image + linguistic)", lambda: ex.explain_caption("synthetic.jpg")),
                 ("explain_caption_wordt + explain_cnn (1 word)", lambda: ex.explain_cnn(ex.explain_caption_wordt(18)[0]))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")

# ---- caption search: the reference-shaped host loop vs the device loop (lrpx.beam), then the full per-image API with
# the model's own caption (encoder once, device beam search, explainer forward, 19+ words explained)
del model.beam_search                                            # back to the class method
wm = ex.word_map
for name, fn in (("beam_search host loop (beam 2, <= 50 steps)", lambda: model.beam_search(img, wm, beam_size=2, max_cap_length=50)),
                 ("beam_search_device (beam 2, 50 steps)", lambda: model.beam_search_device(img, wm, beam_size=2, max_cap_length=50)),
                 ("beam_search host loop (beam 3, <= 20 steps)", lambda: model.beam_search(img, wm, beam_size=3, max_cap_length=20)),
                 ("beam_search_device (beam 3, 20 steps)", lambda: model.beam_search_device(img, wm, beam_size=3, max_cap_length=20))):
    for _ in range(3): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): r = fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms ({len(r[1])} words)")
import contextlib, io
for flag in (True, False):
    ex.DEVICE_BEAM_SEARCH = flag
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(3): ex.explain_caption("synthetic.jpg")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): ex.explain_caption("synthetic.jpg")
        torch.cuda.synchronize()
    print(f"explain_caption incl. caption search, DEVICE_BEAM_SEARCH={flag}: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms "
          f"({ex.caption_length} words)")
