"""Runs in the build container only: the oracle against the UNMODIFIED reference executed live through ref_shim
(fresh seeds, not the committed fixtures).  Exit status 0 = every comparison within the fp32 bar.
Run as a script (tests/test_oracle_vs_reference.py starts it in a clean interpreter: the reference's package names
`models`, `LRPtools` collide with the product's mirrors of the same names)."""
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import lrp_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
import synth  # noqa: E402


def close(a, b, what, rtol=1e-4, atol=1e-6):
    a, b = a.detach().double(), b.detach().double()
    bad = (a - b).abs() > atol + rtol * b.abs()
    if bool(bad.any()):
        raise SystemExit(f"MISMATCH {what}: {int(bad.sum())}/{bad.numel()} worst {float((a - b).abs().max()):.3e}")
    print(f"ok {what}: max abs diff {float((a - b).abs().max()):.3e}")


def main():
    if not ref_shim.reference_available():
        print("reference not mounted")
        return 77
    with contextlib.redirect_stdout(io.StringIO()):
        ns = ref_shim.load_reference()
    # ---- encoder rules: add_lrp + compute_lrp (lrp_wrapper.py:37-87) on a conv/relu/max-pool stack
    cfg = [8, "M", 16, 16, "M", 24]
    net = ns.vgg.make_layers(cfg)
    sd = synth.vgg_state(101, cfg)
    net.load_state_dict(sd)
    net.eval()
    ns.lrp_wrapper.add_lrp(net)
    g = torch.Generator().manual_seed(102)
    x = torch.randn(2, 3, 24, 16, generator=g)
    tgt = torch.randn(2, 24, 6, 4, generator=g)
    rel = net.compute_lrp(x.clone(), target=tgt)
    close(O.sequential_lrp(O.vgg_layers_from_state(sd, cfg), x, tgt), rel, "compute_lrp on conv/relu/pool stack")
    # ---- tuner weights: GridTDModel.get_lrp_weight_step (gridTDmodel.py:549-578)
    V, H, E, B = 90, 48, 16, 5
    stop = synth.stop_mask(V)
    wm = synth.word_map(V)
    rev = {v: ("the" if (bool(stop[v]) and k.startswith("w")) else k) for k, v in wm.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        gm = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    dsd = synth.gridtd_decoder_state(103, V, H, E)
    gm.load_state_dict(dsd, strict=False)
    g = torch.Generator().manual_seed(104)
    logits, h, c = torch.randn(B, V, generator=g), torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    with torch.no_grad():
        wc, wh = gm.get_lrp_weight_step(logits, rev, h, c)
    rc, rh = O.lrp_weight_step(logits, h, c, dsd["fc.weight"], stop)
    close(rc, wc, "get_lrp_weight_step w_ctx", atol=1e-5)
    close(rh, wh, "get_lrp_weight_step w_h", atol=1e-5)
    # ---- ExplainAdaptiveAttention (adaptiveattention.py:626-771) on fresh seeds
    import argparse
    import torch.nn as nn
    V, H, T = 200, 48, 5
    with contextlib.redirect_stdout(io.StringIO()):
        am = ns.adaptiveattention.AdaptiveAttentionCaptioningModel(H, H, V, "vgg16")
    asd = synth.adaptive_decoder_state(105, V, H, H)
    am.load_state_dict(asd, strict=False)
    feats = torch.randn(1, 512, 14, 14, generator=torch.Generator().manual_seed(106)).clamp(min=0)

    class _Enc(nn.Module):
        encoder = nn.Identity()

        def forward(self, img):
            return feats, feats.mean((2, 3)).squeeze()

    am.img_encoder = _Enc()
    toks = synth.tokens(107, T, V)
    am.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
    args = argparse.Namespace(embed_dim=H, hidden_dim=H, encoder="vgg16", height=224, width=224, save_path="/tmp/lrpx_ref",
                              dataset="syn", weight="")
    os.makedirs("/tmp/lrpx_ref", exist_ok=True)
    ex = ns.adaptiveattention.ExplainAdaptiveAttention(args, synth.word_map(V), model=am)
    ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ex.get_hidden_parameters("x")
        rf_ref, rw_ref = ex.explain_caption_wordt(T - 1)
    st = O.adaptive_explainer_forward(asd, feats[0], toks)
    rf, rw, _ = O.adaptive_explain_wordt(asd, st, T - 1)
    close(st["pred"], ex.predictions, "adaptive explainer predictions", atol=2e-5)
    ref = rf_ref[0].reshape(512, -1).t()
    scale = float(ref.abs().max())
    close(rf / scale, ref / scale, "adaptive explain_caption_wordt r_img_feature", rtol=1e-3, atol=1e-5)
    close(rw, rw_ref, "adaptive explain_caption_wordt r_words", rtol=1e-3, atol=1e-5)
    # ---- evaluation.py: block_image (:57-80) and the bbox share (:310-342) on fresh maps
    ev = ref_shim.load_reference_evaluation()

    class _Ex:
        model = nn.Identity()
        word_map = {"<start>": 0}

    exp = ev.EvaluationExperiments(_Ex())
    g = torch.Generator().manual_seed(108)
    heat = torch.randn(1, 3, 64, 96, generator=g)
    with contextlib.redirect_stdout(io.StringIO()):
        mask_ref = exp.block_image(torch.mean(heat, dim=(0, 1)))
    close(O.block_image(heat.mean((0, 1)), exp.num_delete_patches, exp.patch_size), mask_ref, "block_image mask", atol=0)
    import numpy as np
    boxes = [[10, 5, 60, 40], [0, 0, 96, 64], [30, 30, 31, 64]]
    thr = [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    rel = exp._project_maxabs(np.mean(np.maximum(heat.numpy().copy(), 0), axis=(0, 1)))
    want = torch.tensor([[exp._calculate_overlaped_pixels(b, rel, t) for t in thr] for b in boxes])
    close(O.bbox_ratio(heat, boxes, thr), want, "bbox share (in-place quirk)", atol=1e-5)
    return 0


if __name__ == "__main__":
    sys.exit(main())
