"""GPU probe, BASELINE config 3: AOAModelBU on 36 x 2048 bottom-up region features (H = E = 1024, 8 heads, V = 10000,
beam size 3): explanations (region-feature relevance + linguistic relevance of one word) per second through
ExplainAOAAttention.explain_region_features_batch, with and without the caption search, and the per-image API."""
import os, sys, time, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from models import aoamodel as A

dev = "cuda"
V, H, E, B = 10000, 1024, 1024, int(os.environ.get("B", "64"))
model = A.AOAModelBU(E, H, 8, V, "bu")
model.load_state_dict(synth.aoa_bu_state(97, V, H, E), strict=True)
args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="bu", height=224, width=224,
                          save_path="/tmp/lrpx_probe", dataset="syn", weight="")
for prec in ("bf16", "fp32"):
    ex = A.ExplainAOAAttention(args, synth.word_map(V), model=model.to(dev), precision=prec)
    feats = synth.bu_features(98, B).to(dev)
    r = ex.explain_region_features_batch(feats, 3)
    caps = r[4]
    Q = r[0].shape[0]
    toks = torch.tensor([[ex.word_map['<start>']] + c + [0] * (max(map(len, caps)) - len(c)) for c in caps])
    for name, fn in (("beam search + forward + relevance", lambda: ex.explain_region_features_batch(feats, 3)),
                     ("forward + relevance (captions given)", lambda: ex.explain_region_features_batch(feats, 3, tokens=toks))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"config 3 [{prec} decoder GEMMs], {B} feature sets, {Q} (image, word) requests, {name}: {dt * 1e3:.1f} ms = {Q / dt:.0f} explanations/s")
ex.explain_region_features(feats[:1], 3)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): ex.explain_region_features(feats[:1], 3)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"config 3 per-image API (1 feature set, {ex.caption_length} words, beam 3): {dt * 1e3:.1f} ms")
