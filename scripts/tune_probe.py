"""GPU probe, BASELINE config 5: one lrp_tune step (train.py:211-233) on GridTDModel (VGG16 encoder, fixed CNN), batch
128 per GPU, ~20 words, V = 10000 — time per step, and the LRP-weight part alone (get_lrp_weight_step,
gridTDmodel.py:549-578): the batched kernel lrpx_fc_lrp_weights_f32 vs the reference's per-sample loop restated with
tensor ops on the same device (one .item() sync and a V x H temporary per sample, as the reference does)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from models import gridTDmodel as G
from lrpx.tune import LrpTuneStep

dev = "cuda"
B, T, V, H, E = int(os.environ.get("B", "128")), 20, 10000, 512, 512
model = G.GridTDModel(E, H, V, "vgg16")
model.load_state_dict(synth.gridtd_decoder_state(1, V, H, E), strict=False)
model.img_encoder.encoder.load_state_dict(synth.vgg_state(2))
model.to(dev)
wm = synth.word_map(V)
step = LrpTuneStep(model, wm, lr=1e-4, grad_clip=5.0)
imgs = synth.images(3, B).to(dev)
g = torch.Generator().manual_seed(4)
caps = torch.randint(1, V - 4, (B, T + 1), generator=g).to(dev)
caps[:, 0] = wm['<start>']
caplens = [T + 1] * B
for _ in range(2): step.step(imgs, caps, caplens)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 5
for _ in range(n): step.step(imgs, caps, caplens)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print(f"lrp_tune step (config 5), batch {B}, {T} words: {dt * 1e3:.1f} ms per step = {B / dt:.0f} samples/s")

# ---- the LRP weights of one time step
rev = {v: k for k, v in wm.items()}
logits, h, ctx = torch.randn(B, V, device=dev), torch.randn(B, H, device=dev), torch.randn(B, H, device=dev)
for _ in range(3): model.get_lrp_weight_step(logits, rev, h, ctx)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): model.get_lrp_weight_step(logits, rev, h, ctx)
torch.cuda.synchronize(); tk = (time.perf_counter() - t0) / 50
stop = model._stop_mask(rev, logits.device)
W = model.fc.weight.detach()


def per_sample_loop():           # the reference's loop (:552-577) as tensor ops on the device
    wc, wh = torch.ones(B, H, device=dev), torch.ones(B, H, device=dev)
    for b in range(B):
        w = int(torch.argmax(logits[b]).item())
        if bool(stop[w]):
            continue
        r = torch.zeros(1, V, device=dev)
        r[0, w] = logits[b, w]
        s_in = h[b] + ctx[b]
        rel = model.lrp_linear_eps(r.reshape(-1), s_in, logits[b], W)            # V x H temporary inside
        eye = torch.eye(H, device=dev)
        rh = model.lrp_linear_eps(rel, h[b], s_in, eye)
        rc = model.lrp_linear_eps(rel, ctx[b], s_in, eye)
        wh[b] = rh / rh.abs().max().clamp(min=1e-30) + 1
        wc[b] = rc / rc.abs().max().clamp(min=1e-30) + 1
    return wc, wh


per_sample_loop()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): per_sample_loop()
torch.cuda.synchronize(); tl = (time.perf_counter() - t0) / 3
print(f"get_lrp_weight_step, batch {B}: kernel {tk * 1e6:.0f} us vs per-sample loop {tl * 1e3:.1f} ms ({tl / tk:.0f}x)")
