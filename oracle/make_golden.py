"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (through oracle/ref_shim.py).

TEST INFRASTRUCTURE ONLY; run in the build container where /root/reference is mounted:

    python oracle/make_golden.py            # rewrites every fixture

The reference ships no golden vectors, so these fixtures — outputs of the
reference's own code on seeded inputs — are the parity pin for both the oracle restatement
and the CUDA path.  Weights are rebuilt from seeds by tests/synth.py and are not stored.
"""
import argparse
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **_np(arrays))
    print(f"  {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------------------------------
def golden_rules(ns):
    """Single-rule fixtures through lrp_modules.*.propagate_relevance (lrp_modules.py)."""
    lm = ns.lrp_modules
    g = torch.Generator().manual_seed(100)
    out = {}
    # conv alpha-beta, general alpha/beta, stride 2, mixed-sign input, with and without bias
    conv = nn.Conv2d(5, 7, 3, stride=2, padding=1)
    conv.weight.data = torch.randn(7, 5, 3, 3, generator=g) * 0.3
    conv.bias.data = torch.randn(7, generator=g) * 0.1
    x = torch.randn(2, 5, 9, 9, generator=g)
    r = torch.randn(2, 7, 5, 5, generator=g)
    conv.input = (x,)
    for tag, params in [("a1b0", dict(alpha=1., beta=0., ignore_bias=True)),
                        ("a2b1", dict(alpha=2., beta=1., ignore_bias=True)),
                        ("a2b1_bias", dict(alpha=2., beta=1., ignore_bias=False))]:
        ri = lm.Conv2d().propagate_relevance(conv, (torch.zeros_like(x), conv.weight, conv.bias), (r,),
                                             "alpha_beta", params)[0]
        out["conv_R_" + tag] = ri
    out.update(conv_w=conv.weight.data, conv_b=conv.bias.data, conv_x=x, conv_r=r)
    # conv 3x3 s1 p1 on a non-negative input (the VGG case)
    conv2 = nn.Conv2d(6, 4, 3, padding=1)
    conv2.weight.data = torch.randn(4, 6, 3, 3, generator=g) * 0.3
    x2 = torch.randn(1, 6, 8, 8, generator=g).clamp(min=0)
    r2 = torch.randn(1, 4, 8, 8, generator=g)
    conv2.input = (x2,)
    out["conv2_R"] = lm.Conv2d().propagate_relevance(conv2, (torch.zeros_like(x2), conv2.weight, conv2.bias), (r2,),
                                                     "alpha_beta", dict(alpha=1., beta=0., ignore_bias=True))[0]
    out.update(conv2_w=conv2.weight.data, conv2_x=x2, conv2_r=r2)
    # max-pool 2x2 (VGG) with exact-zero windows, 3x3 s2 p1 (ResNet stem, overlapping), avg-pool 2x2
    xp = torch.randn(2, 3, 8, 8, generator=g).clamp(min=0)
    xp[0, 0, :2, :2] = 0
    rp = torch.randn(2, 3, 4, 4, generator=g)
    mp = nn.MaxPool2d(2, 2); mp.input = (xp,)
    out["mp2_R"] = lm.Pool2d().propagate_relevance(mp, None, (rp,), "alpha_beta")[0]
    out["mp2_idx"] = torch.nn.functional.max_pool2d(xp, 2, 2, return_indices=True)[1]
    xq = torch.randn(2, 3, 9, 9, generator=g)
    rq = torch.randn(2, 3, 5, 5, generator=g)
    mp3 = nn.MaxPool2d(3, 2, 1); mp3.input = (xq,)
    out["mp3_R"] = lm.Pool2d().propagate_relevance(mp3, None, (rq,), "alpha_beta")[0]
    out["mp3_idx"] = torch.nn.functional.max_pool2d(xq, 3, 2, 1, return_indices=True)[1]
    ap = nn.AvgPool2d(2, 2); ap.input = (xp,)
    out["ap2_R"] = lm.Pool2d().propagate_relevance(ap, None, (rp,), "alpha_beta")[0]
    out.update(pool_x=xp, pool_r=rp, pool3_x=xq, pool3_r=rq)
    # BatchNorm2d abs-ratio
    bn = nn.BatchNorm2d(3).eval()
    bn.weight.data = torch.rand(3, generator=g) + 0.5; bn.bias.data = torch.randn(3, generator=g) * 0.3
    bn.running_mean = torch.randn(3, generator=g) * 0.3; bn.running_var = torch.rand(3, generator=g) + 0.5
    xb = torch.randn(2, 3, 4, 4, generator=g); xb[0, 0, 0, 0] = 0
    rb = torch.randn(2, 3, 4, 4, generator=g)
    bn.input = (xb,)
    out["bn_R"] = lm.BatchNorm2d().propagate_relevance(bn, (None, None, None), (rb,), "epsilon")[0]
    out.update(bn_x=xb, bn_r=rb, bn_gamma=bn.weight.data, bn_beta=bn.bias.data, bn_mean=bn.running_mean,
               bn_var=bn.running_var, bn_eps=np.float32(bn.eps))
    # Add
    a1 = torch.randn(2, 3, 4, 4, generator=g); a2 = torch.randn(2, 3, 4, 4, generator=g)
    a1[0, 0, 0, :2] = 0; a2[0, 0, 0, :2] = 0          # out == 0 -> 0.5/0.5 split
    ra = torch.randn(2, 3, 4, 4, generator=g)
    add = ns.resnet.Add(); add.input = (a1, a2)
    R1, R2 = lm.Add().propagate_relevance(add, None, (ra,), "alpha_beta")
    out.update(add_x1=a1, add_x2=a2, add_r=ra, add_R1=R1, add_R2=R2)
    # Linear epsilon
    lin = nn.Linear(6, 4)
    lin.weight.data = torch.randn(4, 6, generator=g) * 0.5; lin.bias.data = torch.randn(4, generator=g) * 0.1
    xl = torch.randn(3, 6, generator=g); xl[0, :2] = 0
    rl = torch.randn(3, 4, generator=g)
    for tag, ib in [("nobias", True), ("bias", False)]:
        lin.input = (xl.clone(),)
        out["lin_R_" + tag] = lm.Linear().propagate_relevance(
            lin, (torch.zeros_like(xl), lin.weight.t()), (rl,), "epsilon", dict(ignore_bias=ib))[0]
    out.update(lin_w=lin.weight.data, lin_b=lin.bias.data, lin_x=xl, lin_r=rl)
    # ReLU mask (non-identity) and normalize_relevance
    xr = torch.randn(2, 5, generator=g); rr = torch.randn(2, 5, generator=g)
    relu = nn.ReLU(); relu.input = (xr,)
    out["relu_R_mask"] = lm.ReLU().propagate_relevance(relu, None, (rr,), "mask")[0]
    out.update(relu_x=xr, relu_r=rr)
    xn = torch.randn(4, 6, generator=g); xn[1] = 0
    out.update(norm_x=xn, norm_y=ns.utils.normalize_relevance(xn.clone(), dim=-1))
    save("rules", **out)


def golden_sequential_small(ns):
    """A small conv/relu/maxpool Sequential through add_lrp + compute_lrp (lrp_wrapper.py:37-87)."""
    cfg = [8, 8, "M", 16, 16, "M", 32]
    net = ns.vgg.make_layers(cfg)
    sd = synth.vgg_state(11, cfg)
    net.load_state_dict(sd); net.eval()
    ns.lrp_wrapper.add_lrp(net)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 3, 16, 16, generator=g)
    tgt = torch.randn(2, 32, 4, 4, generator=g)
    rel, logits = net.compute_lrp(x.clone(), target=tgt, return_output=True)
    # Q1: a second call on the same leaf accumulates into sample.grad
    xs = x.clone()
    r1 = net.compute_lrp(xs, target=tgt)
    r2 = net.compute_lrp(xs, target=tgt)
    save("seq_small", cfg=np.array([str(c) for c in cfg]), seed=11, x=x, target=tgt, rel=rel, logits=logits,
         rel_second_call=r2, rel_first_call=r1)


def golden_vgg16(ns):
    for size, seed in [(64, 21), (224, 22)]:
        net = ns.vgg.vgg16(pretrained=True).features[0:-1]      # gridTDmodel.py:33-34
        net.load_state_dict(synth.vgg_state(seed)); net.eval()
        ns.lrp_wrapper.add_lrp(net)
        g = torch.Generator().manual_seed(seed + 1000)
        x = torch.randn(1, 3, size, size, generator=g)
        f = size // 16
        tgt = torch.randn(1, 512, f, f, generator=g) * 1e-3
        rel, feats = net.compute_lrp(x.clone(), target=tgt, return_output=True)
        # the 224 fixture stores only the outputs; x/target are replayed from the seed (see tests)
        extra = dict(x=x, target=tgt, feats=feats) if size == 64 else {}
        save(f"vgg16_{size}", seed=seed, rel=rel, **extra)


def golden_resnet(ns):
    layers = (2, 1, 1, 1)
    net = ns.resnet.ResNet(ns.resnet.Bottleneck, list(layers))
    net.load_state_dict(synth.resnet_state(31, layers)); net.eval()
    ns.lrp_wrapper.add_lrp(net)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 3, 64, 64, generator=g)
    tgt = torch.randn(2, 2048, 2, 2, generator=g)
    rel, feats = net.compute_lrp(x.clone(), target=tgt, return_output=True)
    save("resnet_2111", seed=31, layers=np.array(layers), x=x, target=tgt, rel=rel, feats=feats)


class _StubEncoder(nn.Module):
    """Stands in for models.gridTDmodel.Encoder (gridTDmodel.py:23-43): returns fixed features."""

    def __init__(self, feats):
        super().__init__()
        self.feats = feats
        self.encoder = nn.Identity()

    def forward(self, img):
        return self.feats, self.feats.mean((2, 3)).squeeze()


def _features(seed, C, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, C, h, w, generator=g).clamp(min=0)       # post-ReLU like


def golden_gridtd_decoder(ns):
    for tag, (V, H, E, T, seed, ts) in {"gridtd_dec_512": (1000, 512, 512, 12, 41, [0, 7, 11]),
                                        "gridtd_dec_small": (300, 96, 64, 6, 42, [3])}.items():
        wm = synth.word_map(V)
        with quiet():
            model = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
        sd = synth.gridtd_decoder_state(seed, V, H, E)
        model.load_state_dict(sd, strict=False)
        feats = _features(seed + 1, 512, 14, 14)
        model.img_encoder = _StubEncoder(feats)
        toks = synth.tokens(seed + 2, T, V)
        model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                                  save_path="/tmp/lrpx_ref", dataset="syn", weight="")
        ex = ns.gridTDmodel.ExplainGridTDAttention(args, wm, model=model)
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
        out = dict(V=V, H=H, E=E, T=T, seed=seed, tokens=np.array(toks), ts=np.array(ts), feats=feats,
                   predictions=ex.predictions, alphas=ex.alphas, betas=ex.betas, h2t=ex.h2t, c1t=ex.c1t,
                   context_hat=ex.context_hat)
        for t in ts:
            with torch.no_grad():
                rf, rw = ex.explain_caption_wordt(t)
            out[f"r_feat_{t}"] = rf
            out[f"r_words_{t}"] = rw
        save(tag, **out)


def golden_aoa_decoder(ns):
    cfgs = {"aoa_dec_512": dict(V=1000, H=512, E=512, C=512, hw=(14, 14), T=8, seed=51, cases=[(0, 0), (5, 3)]),
            # config 3 (bottom-up features): 36 regions x 2048-d as a (1,2048,6,6) map, H=1024
            "aoa_dec_bu": dict(V=500, H=1024, E=512, C=2048, hw=(6, 6), T=5, seed=52, cases=[(4, 2)])}
    for tag, c in cfgs.items():
        V, H, E, C, T, seed = c["V"], c["H"], c["E"], c["C"], c["T"], c["seed"]
        wm = synth.word_map(V)
        with quiet():
            model = ns.aoamodel.AOAModel(E, H, 8, V, "vgg16")
        model.img_projector = nn.Conv2d(C, H, 1)
        model.encoder_raw_dim = C
        model.load_state_dict(synth.aoa_decoder_state(seed, V, H, E, C), strict=False)
        feats = _features(seed + 1, C, *c["hw"])
        model.img_encoder = _StubEncoder(feats)
        toks = synth.tokens(seed + 2, T, V)
        model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                                  save_path="/tmp/lrpx_ref", dataset="syn", weight="")
        ex = ns.aoamodel.ExplainAOAAttention(args, wm, model=model)
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
        out = dict(V=V, H=H, E=E, C=C, T=T, seed=seed, tokens=np.array(toks), cases=np.array(c["cases"]), feats=feats,
                   predictions=ex.predictions, alphas=ex.alphas, ht=ex.ht, context_aoa=ex.context_aoa)
        for t, hd in c["cases"]:
            with torch.no_grad():
                rf, rw = ex.explain_caption_wordt(t, hd)
            out[f"r_feat_{t}_{hd}"] = rf
            out[f"r_words_{t}_{hd}"] = rw
        save(tag, **out)


def golden_adaptive_decoder(ns):
    """ExplainAdaptiveAttention (adaptiveattention.py:491-848): saved state + explain_caption_wordt."""
    for tag, (V, H, T, seed, ts) in {"adaptive_dec_512": (1000, 512, 10, 61, [0, 6, 9]),
                                     "adaptive_dec_small": (300, 64, 6, 62, [4])}.items():
        E = H                               # the explainer's x_t buffer is embed_dim + hidden_dim wide (:636)
        wm = synth.word_map(V)
        with quiet():
            model = ns.adaptiveattention.AdaptiveAttentionCaptioningModel(E, H, V, "vgg16")
        model.load_state_dict(synth.adaptive_decoder_state(seed, V, H, E), strict=False)
        feats = _features(seed + 1, 512, 14, 14)
        model.img_encoder = _StubEncoder(feats)
        toks = synth.tokens(seed + 2, T, V)
        model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        args = argparse.Namespace(embed_dim=E, hidden_dim=H, encoder="vgg16", height=224, width=224,
                                  save_path="/tmp/lrpx_ref", dataset="syn", weight="")
        ex = ns.adaptiveattention.ExplainAdaptiveAttention(args, wm, model=model)
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
        out = dict(V=V, H=H, E=E, T=T, seed=seed, tokens=np.array(toks), ts=np.array(ts), feats=feats,
                   predictions=ex.predictions, alphas=ex.alphas, betas=ex.betas, ht=ex.ht, ct=ex.ct, st=ex.st,
                   context_hat=ex.context_hat)
        for t in ts:
            with torch.no_grad(), quiet():
                rf, rw = ex.explain_caption_wordt(t)
            out[f"r_feat_{t}"] = rf
            out[f"r_words_{t}"] = rw
        if tag == "adaptive_dec_small":
            # the model's own beam_search / greedy_search (adaptiveattention.py:370-489) on three feature maps, with the
            # <end> logit raised so that finished and unfinished beams both occur
            del model.beam_search
            end_bias = 0.4
            with torch.no_grad():
                model.fc.bias[V - 1] += end_bias
            beams, greedy = [], []
            for b in range(3):
                model.img_encoder = _StubEncoder(_features(seed + 10 + b, 512, 14, 14))
                for bs in (1, 3):
                    with torch.no_grad(), quiet():
                        _, sen_idx = model.beam_search(torch.zeros(1, 3, 224, 224), wm, beam_size=bs)
                    beams.append(np.array(sen_idx + [-1] * (40 - len(sen_idx)), dtype=np.int64))
                with torch.no_grad(), quiet():
                    _, seqs = model.greedy_search(torch.zeros(1, 3, 224, 224), wm, max_cap_length=12)
                greedy.append(np.array(seqs[0], dtype=np.int64))
            # the model's teacher-forced training forward (adaptiveattention.py:137-190) on two of the feature maps
            fb = torch.cat([_features(seed + 10 + b, 512, 14, 14) for b in range(2)])
            model.img_encoder = _StubEncoder(fb)
            caps = torch.tensor([toks[:5], toks[1:6]])
            with torch.no_grad():
                fpred, falpha, fbeta, _, fmax = model(torch.zeros(2, 3, 224, 224), caps, torch.tensor([5, 4]), None)
            out.update(end_bias=end_bias, beams=np.stack(beams), greedy=np.stack(greedy), fwd_caps=caps,
                       fwd_predictions=fpred, fwd_alphas=falpha, fwd_betas=fbeta, fwd_max_length=np.array(int(fmax)))
        save(tag, **out)


def golden_block_image(ns):
    """EvaluationExperiments.block_image (evaluation.py:57-80): the mask that removes the num_delete_patches most
    relevant 8x8 patches, run by the reference on seeded relevance maps (heat-map-like: channel mean of random maps)."""
    ev = ref_shim.load_reference_evaluation()

    class _Ex:                                            # the two attributes EvaluationExperiments.__init__ touches
        model = nn.Identity()
        word_map = {"<start>": 0}

    exp = ev.EvaluationExperiments(_Ex())
    heat, small = synth.block_image_inputs(71)          # regenerated from the seed by the tests, not stored
    out = dict(seed=np.array(71), k=np.array(exp.num_delete_patches), patch=np.array(exp.patch_size))
    masks = []
    for q in range(heat.shape[0]):
        with quiet():
            masks.append(exp.block_image(torch.mean(heat[q:q + 1], dim=(0, 1))))          # evaluation.py:128-129
    out["masks"] = torch.stack(masks).to(torch.uint8)
    exp.num_delete_patches = 5
    with quiet():
        out["small_masks"] = torch.stack([exp.block_image(torch.mean(small[q:q + 1], dim=(0, 1)))
                                          for q in range(3)]).to(torch.uint8)
    # bounding-box correctness: _project_maxabs + _calculate_overlaped_pixels as bbox_experiment applies them
    # (evaluation.py:403-405, :425-431) on the same heat-maps, three boxes each, ten thresholds
    boxes = synth.bbox_inputs(72, heat.shape[0])
    thresholds = [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    ratios = np.zeros((heat.shape[0], boxes.shape[1], len(thresholds)))
    ratios_neg = np.zeros_like(ratios)
    for q in range(heat.shape[0]):
        for neg, dst in ((False, ratios), (True, ratios_neg)):
            rel = heat[q:q + 1].numpy().copy()
            if neg:
                rel = -1 * rel                                                      # :398-401
            rel = exp._project_maxabs(np.mean(np.maximum(rel, 0), axis=(0, 1)))    # :403-405
            for j, box in enumerate(boxes[q].tolist()):
                for t, thr in enumerate(thresholds):
                    dst[q, j, t] = exp._calculate_overlaped_pixels(box, rel, thr)   # mutates rel like the reference loop
    out.update(bbox_seed=np.array(72), thresholds=np.array(thresholds), ratios=ratios, ratios_neg=ratios_neg)
    save("block_image", **out)


def golden_ablation(ns):
    """EvaluationExperiments.ablation_experiment (evaluation.py:82-290) run END TO END by the reference on a small
    random gridTD / VGG16 model and a seeded 224x224 image: explain_caption (Q1-accumulated heat-maps), then per word the
    image ablation (block_image, beam search on the masked image, teacher-forced score drop or "disappeared") and, from
    word 6 on, the word ablation (3 most relevant preceding words deleted).  The object-word list is widened to the
    synthetic vocabulary (every word counts as a category word) and the caption search is capped at 9 words."""
    ev = ref_shim.load_reference_evaluation()
    V, H, E, seed = 60, 64, 32, 181
    wm = synth.word_map(V)
    with quiet():
        model = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(seed, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(seed + 1))
    model.eval()
    img = synth.images(seed + 2, 1)
    orig_search = model.beam_search
    model.beam_search = lambda im, w, beam_size=3, max_cap_length=20: orig_search(im, w, beam_size=beam_size, max_cap_length=9)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path="/tmp/lrpx_ref", dataset="syn", weight="")
    ex = ns.gridTDmodel.ExplainGridTDAttention(args, wm, model=model)
    ex.preprocess_img = lambda p: img.clone()
    ex.visualize_explanations = lambda *a, **k: None
    ev.object_words_list = [w for w in wm if w.startswith("w")]
    exp = ev.EvaluationExperiments(ex)
    os.makedirs("/tmp/lrpx_ref/abl", exist_ok=True)
    with quiet():
        exp.ablation_experiment({"image_path": "/tmp/lrpx_ref/synthetic.jpg"}, "lrp", "/tmp/lrpx_ref/abl", do_attention=False)
    toks = ex.beam_caption_encode
    T = len(toks) - 1
    disappear = np.zeros(T, dtype=np.int64)
    img_diff = np.full(T, np.nan)
    for t, w in exp.image_disappear_count:
        disappear[int(t)] = 1
    for t, w, d in exp.image_category_score_diff:
        img_diff[int(t)] = d
    word_diff = np.full(T, np.nan)
    for t, ds in exp.category_scores_diff.items():
        word_diff[int(t)] = ds[0]
    assert not exp.stop_word_scores_diff
    # tpfp_experiment (evaluation.py:450-573) on the same image: every synthetic word is "frequent"; the reference
    # captions contain the caption's last word but not its first, so both a TP and an FP branch are taken
    ref_caps = [[wm["<start>"], toks[-1], 7, wm["<end>"]], [wm["<start>"], 9, toks[-1], wm["<end>"], wm["<pad>"]]]
    with quiet():
        exp.tpfp_experiment({"image_path": "/tmp/lrpx_ref/synthetic.jpg", "encoded_all_caps": [list(c) for c in ref_caps]},
                            "lrp", "/tmp/lrpx_ref/abl", [w for w in wm if w.startswith("w")], do_attention=False)
    assert ex.beam_caption_encode == toks
    stat = lambda rows, k: np.array([float(r[k]) for r in rows])
    quant = lambda rows: np.array([[float(v) for v in r["quantile"]] for r in rows])
    tp, fp = exp.TP_statistics, exp.FP_statistics
    tpfp = dict(ref_caps_0=np.array(ref_caps[0]), ref_caps_1=np.array(ref_caps[1]),
                tp_words=np.array([wm[r["word"]] for r in tp]), fp_words=np.array([wm[r["word"]] for r in fp]),
                tp_beta=np.array([float(r["1-beta"]) for r in exp.TP_statistics_beta]),
                fp_beta=np.array([float(r["1-beta"]) for r in exp.FP_statistics_beta]))
    for k in ("mean", "mean_abs", "mean_pos", "max"):
        tpfp["tp_" + k], tpfp["fp_" + k] = stat(tp, k), stat(fp, k)
    tpfp["tp_quantile"], tpfp["fp_quantile"] = quant(tp), quant(fp)
    save("ablation_e2e", seed=np.array(seed), V=np.array(V), H=np.array(H), E=np.array(E), tokens=np.array(toks),
         disappear=disappear, img_diff=img_diff, word_diff=word_diff, predictions=ex.predictions.detach(), **tpfp)
    print("tpfp:", len(tp), "TP words,", len(fp), "FP words")
    print("ablation fixture:", toks, disappear.tolist(), img_diff.tolist(), word_diff.tolist())


def _ref_gradient_explainer(ns_mod, cls_name, model, args, wm, tag):
    """Builds one of the reference's gradient-family explainers through its own constructor (which loads a checkpoint
    from args.weight), then swaps in the seeded model."""
    path = f"/tmp/lrpx_ref/{tag}.pth"
    torch.save({"state_dict": model.state_dict()}, path)
    args = argparse.Namespace(**{**vars(args), "weight": path})
    with quiet():
        ex = getattr(ns_mod, cls_name)(args, wm)
    return ex


def golden_gradient(ns):
    """Gradient-family explainers, run by the reference itself.
    gridtd_grad_512 / aoa_grad_512: the hand-written decoder backward at the benched decoder size over fixed features
    (stub encoder): saved gates, d_img_feature and r_words of ExplainGridTDGradient / ExplainiGridTDGuidedGradient /
    ExplainAOAGradient, and the Grad-CAM maps.
    gradient_e2e: a small decoder on a seeded VGG16 at 224x224: explain_cnn of the plain gradient (autograd) and of
    guided backpropagation (the reference's ReLU hooks).  Guided Grad-CAM needs skimage (absent here): unpinned."""
    # ---------------------------------------------------------------- gridTD decoder
    V, H, E, T, seed, ts = 1000, 512, 512, 12, 201, [0, 7, 11]
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path="/tmp/lrpx_ref", dataset="syn", weight="")
    with quiet():
        model = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(seed, V, H, E), strict=False)
    feats = _features(seed + 1, 512, 14, 14)
    toks = synth.tokens(seed + 2, T, V)
    out = dict(V=V, H=H, E=E, T=T, seed=seed, tokens=np.array(toks), ts=np.array(ts), feats=feats)
    for cls, key in (("ExplainGridTDGradient", "grad"), ("ExplainiGridTDGuidedGradient", "guided"),
                     ("ExplainGridTDGradCam", "cam")):
        ex = _ref_gradient_explainer(ns.gridTDmodel, cls, model, args, wm, "gridtd_grad")
        ex.model.img_encoder = _StubEncoder(feats)
        ex.model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
            ex.image_feature_proj = ex.image_feature_proj.transpose(1, 2)          # as explain_caption does (:1528)
        if key == "grad":
            out.update(predictions=ex.predictions, o1=ex.o1t_act, o2=ex.o2t_act, sen_gate=ex.sen_gate, betas=ex.betas,
                       c1t=ex.c1t, c2t=ex.c2t)
        for t in ts:
            with torch.no_grad():
                df, rw = ex.explain_caption_wordt(t)
                if key == "cam":
                    out[f"cam_{t}"] = ex.explain_cnn(df)
                else:
                    out[f"{key}_d_feat_{t}"] = df
                    out[f"{key}_r_words_{t}"] = rw
    save("gridtd_grad_512", **out)
    # ---------------------------------------------------------------- AoA decoder
    V, H, E, C, T, seed, cases = 1000, 512, 512, 512, 8, 211, [(0, 0), (5, 3), (7, 7)]
    wm = synth.word_map(V)
    with quiet():
        model = ns.aoamodel.AOAModel(E, H, 8, V, "vgg16")
    model.load_state_dict(synth.aoa_decoder_state(seed, V, H, E, C), strict=False)
    feats = _features(seed + 1, C, 14, 14)
    toks = synth.tokens(seed + 2, T, V)
    out = dict(V=V, H=H, E=E, C=C, T=T, seed=seed, tokens=np.array(toks), cases=np.array(cases), feats=feats)
    for cls, key in (("ExplainAOAGradient", "grad"), ("ExplainAOAGradCam", "cam")):
        ex = _ref_gradient_explainer(ns.aoamodel, cls, model, args, wm, "aoa_grad")
        ex.model.img_encoder = _StubEncoder(feats)
        ex.model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
        if key == "grad":
            out.update(predictions=ex.predictions, ot=ex.ot_act, context_aoa_gate=ex.context_aoa_gate)
        for t, hd in cases:
            with torch.no_grad():
                df, rw = ex.explain_caption_wordt(t, hd)
                if key == "cam":
                    out[f"cam_{t}_{hd}"] = ex.explain_cnn(df)
                else:
                    out[f"d_feat_{t}_{hd}"] = df
                    out[f"r_words_{t}_{hd}"] = rw
    save("aoa_grad_512", **out)
    # ---------------------------------------------------------------- end to end through a seeded VGG16
    V, H, E, seed = 60, 64, 32, 221
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path="/tmp/lrpx_ref", dataset="syn", weight="")
    with quiet():
        model = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(seed, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(seed + 1))
    img = synth.images(seed + 2, 1)
    out = dict(V=V, H=H, E=E, seed=seed)
    toks = None
    for cls, key in (("ExplainGridTDGradient", "grad"), ("ExplainiGridTDGuidedGradient", "guided")):
        ex = _ref_gradient_explainer(ns.gridTDmodel, cls, model, args, wm, "grad_e2e")
        orig = ex.model.beam_search
        ex.model.beam_search = lambda im, w, beam_size=3, max_cap_length=50, orig=orig: orig(im, w, beam_size=beam_size, max_cap_length=7)
        ex.preprocess_img = lambda p: img.clone()
        with quiet():
            imgs, words = ex.explain_caption("x")
        if toks is None:
            toks = list(ex.beam_caption_encode)
            out["tokens"] = np.array(toks)
        assert list(ex.beam_caption_encode) == toks
        T = len(toks) - 1
        for t in (0, T - 1):
            out[f"{key}_img_{t}"] = imgs[t].detach()
        out[f"{key}_words"] = np.concatenate([w.detach().numpy() for w in words])
    save("gradient_e2e", **out)
    print("gradient_e2e tokens:", toks)


def golden_gradient_adaptive(ns):
    """ExplainAdaptiveGradient / ExplainiAdaptiveGuidedGradient (adaptiveattention.py:851-1216) run by the reference over
    fixed features (stub encoder): saved gates, d_img_feature and r_words."""
    V, H, T, seed, ts = 1000, 512, 10, 241, [0, 6, 9]
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=H, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path="/tmp/lrpx_ref", dataset="syn", weight="")
    with quiet():
        model = ns.adaptiveattention.AdaptiveAttentionCaptioningModel(H, H, V, "vgg16")
    model.load_state_dict(synth.adaptive_decoder_state(seed, V, H, H), strict=False)
    feats = _features(seed + 1, 512, 14, 14)
    toks = synth.tokens(seed + 2, T, V)
    out = dict(V=V, H=H, T=T, seed=seed, tokens=np.array(toks), ts=np.array(ts), feats=feats)
    for cls, key in (("ExplainAdaptiveGradient", "grad"), ("ExplainiAdaptiveGuidedGradient", "guided")):
        ex = _ref_gradient_explainer(ns.adaptiveattention, cls, model, args, wm, "adaptive_grad")
        ex.model.img_encoder = _StubEncoder(feats)
        ex.model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in toks[1:])], toks[1:])
        ex.preprocess_img = lambda p: torch.zeros(1, 3, 224, 224)
        with torch.no_grad(), quiet():
            ex.get_hidden_parameters("x")
            ex.image_feature_proj = ex.image_feature_proj.transpose(1, 2)
        if key == "grad":
            out.update(predictions=ex.predictions, ot=ex.ot_act, sen_gate=ex.sen_gate, betas=ex.betas)
        for t in (ts if key == "grad" else ts[-1:]):         # the guided decoder half has the same values: one word suffices
            with torch.no_grad():
                df, rw = ex.explain_caption_wordt(t)
            out[f"{key}_d_feat_{t}"] = df
            out[f"{key}_r_words_{t}"] = rw
    save("adaptive_grad_512", **out)


def golden_scst(ns):
    """get_self_critical_reward (models/modelutils.py:200-238) run by the reference with its vendored CIDEr / BLEU
    scorers on seeded sampled / greedy / ground-truth captions: CIDEr-only (the training setting, train.py:193), BLEU-only
    and a mix; plus RewardCriterion."""
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_modelutils", os.path.join(ref_shim.REFERENCE_ROOT, "models/modelutils.py"))
        mu = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mu)
    finally:
        sys.path.remove(ref_shim.REFERENCE_ROOT)
    V, B, L, R = 40, 12, 9, 3
    wm = synth.word_map(V)
    g = torch.Generator().manual_seed(231)

    def caps(n, length):
        out = torch.randint(1, 14, (n, length), generator=g)              # a small active vocabulary: n-grams repeat
        for i in range(n):
            k = int(torch.randint(3, length + 1, (1,), generator=g))
            if k < length:
                out[i, k] = wm['<end>']
                out[i, k + 1:] = 0
        return out
    gen, greedy = caps(B, L), caps(B, L)
    greedy[0] = gen[0]
    gts = caps(B * R, L + 2).view(B, R, L + 2)
    gts[:, :, 0] = wm['<start>']
    gts[1, 0, 1:L + 1] = gen[1]                                            # an exact match among the references
    data_gts = [[gts[b, r].numpy() for r in range(R)] for b in range(B)]
    out = dict(V=V, gen=gen, greedy=greedy, gts=gts)
    for tag, (cw, bw) in {"cider": (1.0, 0.0), "bleu": (0.0, 1.0), "mix": (0.7, 0.3)}.items():
        out["reward_" + tag] = mu.get_self_critical_reward(greedy, data_gts, gen, wm, cw, bw)
    logp = -torch.rand(B, L, generator=g)
    rew = torch.from_numpy(out["reward_cider"]).float()
    out["logp"] = logp
    out["loss"] = mu.RewardCriterion()(logp, gen, rew)
    save("scst_reward", **out)


def _rev_word_map(V, stop):
    wm = synth.word_map(V)
    rev = {v: k for k, v in wm.items()}
    for i in range(V):
        if stop[i] and rev[i].startswith("w"):
            rev[i] = "the"                      # a member of the shim's stop-word list
    return wm, rev


def golden_lrp_weights(ns):
    V, H, E, B = 200, 64, 32, 8
    stop = synth.stop_mask(V)
    wm, rev = _rev_word_map(V, stop)
    with quiet():
        gm = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    sd = synth.gridtd_decoder_state(61, V, H, E)
    gm.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(62)
    logits = torch.randn(B, V, generator=g); h = torch.randn(B, H, generator=g); c = torch.randn(B, H, generator=g)
    logits[0, 7] = 9.0          # stop word (7 % 7 == 0)
    logits[1, V - 1] = 9.0      # <end>
    h[2] = 0; c[2] = 0          # all-zero relevance row -> weights 1
    with torch.no_grad():
        wc, wh = gm.get_lrp_weight_step(logits, rev, h, c)
    save("lrp_weights", V=V, H=H, E=E, seed=61, logits=logits, h=h, ctx=c, w_ctx=wc, w_h=wh, stop=stop)


def golden_tune(ns):
    """forwardlrp_context (gridTDmodel.py:580-633) on the full model incl. the VGG16 encoder."""
    V, H, E, B, L = 60, 32, 32, 2, 5
    stop = synth.stop_mask(V)
    wm, rev = _rev_word_map(V, stop)
    with quiet():
        gm = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
    gm.load_state_dict(synth.gridtd_decoder_state(71, V, H, E), strict=False)
    gm.img_encoder.encoder.load_state_dict(synth.vgg_state(72))
    gm.eval()
    imgs = synth.images(73, B)
    g = torch.Generator().manual_seed(74)
    caps = torch.randint(1, V - 4, (B, L), generator=g)
    caps[:, 0] = V - 2
    caplens = torch.tensor([L, L - 1])
    with torch.no_grad():
        pred, wpred, maxlen = gm.forwardlrp_context(imgs, caps, caplens, rev)
        # the greedy baseline of trainciderlrp (train.py:259-261): GridTDModel.sample (:200-242)
        seq, seq_lp, _ = gm.sample(imgs, wm, torch.tensor([9, 9]))
    save("tune_gridtd", V=V, H=H, E=E, seeds=np.array([71, 72, 73, 74]), caps=caps, caplens=caplens, stop=stop,
         predictions=pred, weighted_predictions=wpred, max_length=int(maxlen), sample_seq=seq, sample_logprobs=seq_lp)


END_BIAS = float(os.environ.get("END_BIAS", "0.17"))   # added to the <end> logit bias of the BU fixtures


def golden_tune_bu(ns):
    """Bottom-up twins (gridTDmodel.py:1863-2478, aoamodel.py:1779-2400) on 36 x 2048 region features:
    forwardlrp_context, sample_lrp (greedy) and beam_search(beam_size=3) of the reference itself.  The <end> logit
    gets a positive bias so that beams do finish (complete / incomplete bookkeeping, bit-exact beam indices)."""
    V, H, E, B, L = 60, 64, 32, 3, 6
    stop = synth.stop_mask(V)
    wm, rev = _rev_word_map(V, stop)
    feats = synth.bu_features(83, B)
    g = torch.Generator().manual_seed(84)
    caps = torch.randint(1, V - 4, (B, L), generator=g)
    caps[:, 0] = V - 2
    caplens = torch.tensor([L, L - 1, L])
    out = dict(V=V, H=H, E=E, end_bias=END_BIAS, seeds=np.array([81, 82, 83, 84]), caps=caps, caplens=caplens, stop=stop)
    for tag, make, state in (("gridtd", lambda: ns.gridTDmodel.GridTDModelBU(E, H, V, "bu"), synth.gridtd_bu_state(81, V, H, E)),
                             ("aoa", lambda: ns.aoamodel.AOAModelBU(E, H, 8, V, "bu"), synth.aoa_bu_state(82, V, H, E))):
        with quiet():
            m = make()
        state["fc.bias"][V - 1] += END_BIAS
        m.load_state_dict(state, strict=True)
        m.eval()
        with torch.no_grad():
            pred, wpred, maxlen = m.forwardlrp_context(feats, caps, caplens, rev)
            seq, seq_lp, _ = m.sample_lrp(feats, rev, wm, caplens, {"sample_method": "greedy"})
            beams = []
            for b in range(B):
                for bs in (1, 3):
                    _, sen_idx = m.beam_search(feats[b:b + 1], wm, beam_size=bs)
                    beams.append(np.array(sen_idx + [-1] * (40 - len(sen_idx)), dtype=np.int64))
            # plain teacher-forced forward and greedy sample of the twins (gridTDmodel.py:1899-1955,:1956-1998 /
            # aoamodel.py:1826-1881,:1883-1925)
            fwd = m(feats, caps, caplens, None)
            sseq, sseq_lp, _ = m.sample(feats, wm, caplens)
        out.update({f"{tag}_predictions": pred, f"{tag}_weighted_predictions": wpred, f"{tag}_max_length": int(maxlen),
                    f"{tag}_seq": seq, f"{tag}_seq_logprobs": seq_lp, f"{tag}_beams": np.stack(beams),
                    f"{tag}_fwd_predictions": fwd[0], f"{tag}_sample_seq": sseq, f"{tag}_sample_logprobs": sseq_lp})
    save("tune_bu", **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    os.makedirs("/tmp/lrpx_ref", exist_ok=True)
    ns = ref_shim.load_reference()
    torch.manual_seed(0)
    only = set(sys.argv[1:])          # optional: names of the generators to (re)run
    for fn in (golden_rules, golden_sequential_small, golden_vgg16, golden_resnet, golden_gridtd_decoder,
               golden_aoa_decoder, golden_adaptive_decoder, golden_block_image, golden_lrp_weights, golden_tune, golden_tune_bu,
               golden_ablation, golden_gradient, golden_gradient_adaptive, golden_scst):
        if only and fn.__name__ not in only:
            continue
        print(fn.__name__)
        fn(ns)


if __name__ == "__main__":
    main()
