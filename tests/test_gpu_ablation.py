"""GPU: patch-ablation masks (lrpx_block_image_f32) against masks produced by the reference's own
EvaluationExperiments.block_image (fixture block_image), and the batched ablation experiments (lrpx.ablation) against
the per-request walk of evaluation.py:120-150 / :234-262 done with the mirror's host functions."""
import argparse

import pytest
import torch

import lrp_oracle as O
import synth
from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_block_image_masks_bit_exact_vs_reference(golden):
    from lrpx import ops
    g = golden("block_image")
    heat, small = synth.block_image_inputs(int(g["seed"]))
    k, patch = int(g["k"]), int(g["patch"])
    mask = ops.block_image(heat.to(DEV), k, patch)
    assert torch.equal(mask.cpu().to(torch.uint8), g["masks"])
    assert torch.equal(ops.block_image(small.to(DEV), 5, patch).cpu().to(torch.uint8), g["small_masks"])
    # masked images: mask * image of the request's image (evaluation.py:130), requests in arbitrary order
    imgs = synth.images(5, 3).to(DEV)
    req = torch.tensor([2, 0, 0, 1, 2, 1], dtype=torch.int32)
    m2, masked = ops.block_image(heat.to(DEV), k, patch, images=imgs, req_img=req)
    assert torch.equal(m2, mask)
    assert torch.equal(masked, mask.unsqueeze(1) * imgs[req.long().to(DEV)])
    # edge cases: k = 0 keeps everything, k = all patches blanks everything, empty batch
    assert bool((ops.block_image(small.to(DEV), 0, patch) == 1).all())
    assert bool((ops.block_image(small.to(DEV), 24, patch) == 0).all())
    assert ops.block_image(small[:0].to(DEV), 5, patch).shape == (0, 32, 48)
    from lrpx._lib import LrpxError
    with pytest.raises(LrpxError):
        ops.block_image(small.to(DEV), 25, patch)                  # more patches than there are (evaluation.py:65)
    with pytest.raises(LrpxError):
        ops.block_image(small[:, :, :30].contiguous().to(DEV), 5, patch)   # H not a multiple of the patch size (:59)


def test_bbox_ratio_vs_reference(golden):
    """lrpx_bbox_ratio_f32 against the values the reference's _project_maxabs / _calculate_overlaped_pixels returned
    (fixture block_image), positive and negative relevance, ragged box counts."""
    from lrpx import ops
    g = golden("block_image")
    heat, _ = synth.block_image_inputs(int(g["seed"]))
    boxes = synth.bbox_inputs(int(g["bbox_seed"]), heat.shape[0])
    thr = g["thresholds"].tolist()
    got = ops.bbox_ratio(heat.to(DEV), boxes, thresholds=thr)
    neg = ops.bbox_ratio(heat.to(DEV), boxes, thresholds=thr, negative=True)
    assert float((got.cpu().double() - g["ratios"]).abs().max()) <= 1e-5
    assert float((neg.cpu().double() - g["ratios_neg"]).abs().max()) <= 1e-5
    assert float(got[0, -1].abs().max()) == 0.0                       # the empty box
    nb = torch.tensor([3, 1, 2, 0, 3, 2], dtype=torch.int32)
    rag = ops.bbox_ratio(heat.to(DEV), boxes, n_boxes=nb, thresholds=thr)
    for q in range(6):
        assert torch.equal(rag[q, :int(nb[q])], got[q, :int(nb[q])]) and float(rag[q, int(nb[q]):].abs().sum()) == 0.0
    # every pair on the fresh map (inplace_quirk=False): box 0 as in the reference, the others vs the oracle
    fresh = ops.bbox_ratio(heat.to(DEV), boxes, thresholds=thr, inplace_quirk=False)
    for q in range(6):
        want = O.bbox_ratio(heat[q:q + 1].double(), boxes[q].tolist(), thr, inplace_quirk=False)
        assert float((fresh[q].cpu().double() - want).abs().max()) <= 1e-5
    assert torch.equal(fresh[:, 0], got[:, 0])
    zero = ops.bbox_ratio(torch.zeros(1, 3, 224, 224, device=DEV), boxes[:1], thresholds=thr)
    assert float(zero.abs().max()) == 0.0                              # total relevance 0 -> 0 (:328-329)


def test_tpfp_statistics_vs_numpy():
    """evaluation.py:503-513 per request with numpy (the reference's own calls) vs the batched device form."""
    import numpy as np
    from lrpx.ablation import AblationExperiments
    heat, _ = synth.block_image_inputs(71)
    heat[3] = -heat[3].abs()                                            # a map without positive pixels: mean_pos = 0
    st = AblationExperiments.tpfp_statistics(heat.to(DEV))
    qp = [i / 100 for i in range(0, 100)]
    for q in range(heat.shape[0]):
        r = np.mean(heat[q:q + 1].numpy(), axis=(0, 1))
        mean_pos = 0 if np.sum(r > 0) == 0 else np.sum(np.maximum(r, 0)) / np.sum(r > 0)
        for key, want in (("mean", np.mean(r)), ("mean_abs", np.mean(np.abs(r))), ("mean_pos", mean_pos), ("max", np.max(r))):
            assert abs(float(st[key][q]) - float(want)) <= 1e-5 * abs(float(want)) + 1e-7, (q, key)
        assert np.allclose(st["quantile"][q].cpu().numpy(), np.quantile(r, qp), rtol=1e-5, atol=1e-7), q


def test_batched_ablation_equals_per_request_walk(tmp_path):
    """B images x T words on GridTDModel + VGG16 (bf16 chain): explain, then the image ablation and the word ablation
    for all requests at once == one request at a time with the mirror's host-side beam_search / teacherforce_forward
    over the same (tensor-core) features and the oracle's block_image."""
    from models import gridTDmodel as G
    from lrpx.pipeline import BatchExplainer
    from lrpx.ablation import AblationExperiments
    V, H, E, B, T = 60, 64, 32, 2, 5
    model = G.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(501, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(502))
    model.to(DEV).eval()
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path=str(tmp_path), dataset="syn", weight="")
    ex = G.ExplainGridTDAttention(args, wm, model=model, precision="bf16")
    imgs = synth.images(503, B).to(DEV)
    toks = torch.stack([torch.tensor(synth.tokens(504 + b, T, V)) for b in range(B)]).to(DEV)
    heat, r_words = BatchExplainer(ex, chunk=8).explain(imgs, toks)
    req_img = torch.arange(B, dtype=torch.int32, device=DEV).repeat_interleave(T)
    req_t = torch.arange(T, dtype=torch.int32, device=DEV).repeat(B)
    eng = ex.engine()
    feat = eng.features(eng.forward(imgs), "pixel").clone()
    pred = ex.explainer_forward(feat, toks)["pred"]
    ab = AblationExperiments(ex, chunk=4)
    out = ab.image_ablation(imgs, toks, heat, req_img, req_t, pred)
    # ---- per-request walk (evaluation.py:120-150)
    n_gone = 0
    for q in range(B * T):
        b, t = q // T, q % T
        word = int(toks[b, t + 1])
        mask = O.block_image(heat[q:q + 1].mean((0, 1)).cpu(), 20, 8).to(DEV)
        assert torch.equal(ab.block_image(heat[q:q + 1].mean((0, 1))), mask)
        image_modified = mask * imgs[b:b + 1]
        assert torch.equal(out["masked"][q:q + 1], image_modified)
        fmap = eng.features(eng.forward(image_modified), "pixel")[0].t().reshape(1, 512, 14, 14).clone()

        class Stub(torch.nn.Module):
            encoder = torch.nn.Identity()

            def forward(self, im):
                return fmap, fmap.mean((2, 3)).squeeze()

        real = model.img_encoder
        model.img_encoder = Stub()
        _, new_idx = model.beam_search(image_modified, wm)
        model.img_encoder = real
        assert out["captions"][q] == new_idx, (q, out["captions"][q], new_idx)
        if word in new_idx:
            prefix = [wm['<start>']] + new_idx[:new_idx.index(word)]
            new_scores = ex.teacherforce_forward(image_modified, prefix)
            want = torch.softmax(pred[b, t], -1)[word] - torch.softmax(new_scores[-1], -1)[word]
            assert not bool(out["disappear"][q])
            assert abs(float(out["score_diff"][q]) - float(want)) <= 1e-3 * abs(float(want)) + 1e-6, (q, out["score_diff"][q], want)
        else:
            n_gone += 1
            assert bool(out["disappear"][q]) and bool(torch.isnan(out["score_diff"][q]))
    print(f"image ablation: {n_gone} of {B * T} words disappeared")
    # ---- word ablation (evaluation.py:234-262) for the requests with t >= 3
    sel = [q for q in range(B * T) if q % T >= 3]
    idx = torch.tensor(sel, device=DEV)
    diff = ab.word_ablation(feat, toks, r_words[idx], req_img[idx], req_t[idx], pred)
    for n, q in enumerate(sel):
        b, t = q // T, q % T
        word = int(toks[b, t + 1])
        top = torch.topk(r_words[q, 1:t + 1], k=3).indices.cpu().numpy()
        import numpy as np
        deleted = list(np.delete(np.array(toks[b, :t + 1].tolist()), top + 1))
        new_scores = ex.teacherforce_forward(imgs[b:b + 1], [int(v) for v in deleted])
        want = torch.softmax(pred[b, t], -1)[word] - torch.softmax(new_scores[-1], -1)[word]
        assert abs(float(diff[n]) - float(want)) <= 1e-3 * abs(float(want)) + 1e-6, (q, diff[n], want)


def test_ablation_experiment_end_to_end_vs_reference_fixture(golden, tmp_path):
    """evaluation.py:82-290 END TO END against the reference itself (fixture ablation_e2e: the reference's own
    ablation_experiment on a seeded random gridTD / VGG16 model, caption search capped at 9 words, every word a category
    word): explain_caption with the reference's Q1 accumulation, image ablation of the words t >= 1 (mask of the 20 most
    relevant patches of the ACCUMULATED heat-map, beam search on the masked image, teacher-forced score drop) and word
    ablation of the words t >= 6.  fp32-accurate mode; the caption, the disappear flags and the score drops must come
    out as the reference's (scores are softmax values ~1/60: drops of ~3e-4 compared at 5e-6 absolute)."""
    import numpy as np
    from models import gridTDmodel as G
    from lrpx.ablation import AblationExperiments
    g = golden("ablation_e2e")
    V, H, E, seed = int(g["V"]), int(g["H"]), int(g["E"]), int(g["seed"])
    model = G.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(seed, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(seed + 1))
    model.to(DEV).eval()
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path=str(tmp_path), dataset="syn", weight="")
    ex = G.ExplainGridTDAttention(args, wm, model=model, precision="fp32")
    img = synth.images(seed + 2, 1).to(DEV)
    ex.preprocess_img = lambda p: img
    # the explainer's own caption search is beam 2 / max 50 (gridTDmodel.py:935); the fixture capped the reference at 9
    find = ex._find_caption
    ex._find_caption = lambda path, beam_size, max_cap_length: find(path, beam_size, 9)
    heat_l, words_l = ex.explain_caption("synthetic.jpg")               # Q1: running sums, like the reference
    toks = ex.beam_caption_encode
    assert toks == g["tokens"].tolist(), (toks, g["tokens"].tolist())
    T = len(toks) - 1
    assert_close(ex.predictions, g["predictions"], rtol=1e-3, atol=1e-4, what="explainer predictions")
    heat = torch.cat(heat_l)
    tokens = torch.tensor([toks], device=DEV)
    pred = ex.predictions.unsqueeze(0)
    ab = AblationExperiments(ex)
    ts = list(range(1, T))
    req_img = torch.zeros(len(ts), dtype=torch.int32, device=DEV)
    req_t = torch.tensor(ts, dtype=torch.int32, device=DEV)
    out = ab.image_ablation(img, tokens, heat[1:], req_img, req_t, pred, beam_size=3, max_cap_length=9)
    want_gone, want_diff = g["disappear"].numpy(), g["img_diff"].numpy()
    for n, t in enumerate(ts):
        assert bool(out["disappear"][n]) == bool(want_gone[t]), (t, out["captions"][n])
        if not want_gone[t]:
            assert abs(float(out["score_diff"][n]) - want_diff[t]) <= 5e-6, (t, float(out["score_diff"][n]), want_diff[t])
    # word ablation from word 6 on
    ts6 = [t for t in range(6, T)]
    r_words = torch.zeros(len(ts6), T, device=DEV)
    for n, t in enumerate(ts6):
        r_words[n, :t + 1] = words_l[t]
    feat = ex._state["feat"]
    diff = ab.word_ablation(feat, tokens, r_words, torch.zeros(len(ts6), dtype=torch.int32, device=DEV),
                            torch.tensor(ts6, dtype=torch.int32, device=DEV), pred)
    want_w = g["word_diff"].numpy()
    for n, t in enumerate(ts6):
        assert abs(float(diff[n]) - want_w[t]) <= 5e-6, (t, float(diff[n]), want_w[t])
    print("image ablation drops", [float(v) for v in out["score_diff"]], "word ablation drops", [float(v) for v in diff])


def test_tpfp_experiment_end_to_end_vs_reference_fixture(golden, tmp_path):
    """evaluation.py:450-573 END TO END against the reference (same fixture): which caption words are true / false
    positives against the reference captions, and per word the statistics of the channel-mean of its (Q1-accumulated)
    heat-map — mean, mean |.|, mean of the positive part, max, the 100 quantiles — plus 1 - beta of the sentinel gate.
    fp32-accurate chain: statistics within 2e-3 relative (+1e-4 of the largest quantile)."""
    from models import gridTDmodel as G
    from lrpx.ablation import AblationExperiments
    g = golden("ablation_e2e")
    V, H, E, seed = int(g["V"]), int(g["H"]), int(g["E"]), int(g["seed"])
    model = G.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(seed, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(seed + 1))
    model.to(DEV).eval()
    wm = synth.word_map(V)
    args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                              save_path=str(tmp_path), dataset="syn", weight="")
    ex = G.ExplainGridTDAttention(args, wm, model=model, precision="fp32")
    img = synth.images(seed + 2, 1).to(DEV)
    ex.preprocess_img = lambda p: img
    find = ex._find_caption
    ex._find_caption = lambda path, beam_size, max_cap_length: find(path, beam_size, 9)
    heat_l, _ = ex.explain_caption("synthetic.jpg")
    toks = ex.beam_caption_encode
    assert toks == g["tokens"].tolist()
    frequent = [i for w, i in wm.items() if w.startswith("w")]
    special = [wm[k] for k in ("<start>", "<pad>", "<end>", "<unk>")]
    tp, fp = AblationExperiments.tpfp_split(toks, frequent, [g["ref_caps_0"].tolist(), g["ref_caps_1"].tolist()], special)
    assert [toks[t + 1] for t in tp] == g["tp_words"].tolist() and [toks[t + 1] for t in fp] == g["fp_words"].tolist()
    assert tp and fp
    heat = torch.cat(heat_l)
    for name, ts in (("tp", tp), ("fp", fp)):
        st = AblationExperiments.tpfp_statistics(heat[torch.tensor(ts, device=DEV)])
        scale = float(g[name + "_quantile"].abs().max())
        for k in ("mean", "mean_abs", "mean_pos", "max"):
            assert_close(st[k], g[name + "_" + k], rtol=2e-3, atol=1e-4 * scale, what=f"{name} {k}")
        assert_close(st["quantile"], g[name + "_quantile"], rtol=2e-3, atol=1e-4 * scale, what=f"{name} quantiles")
        one_minus_beta = 1 - ex.betas[torch.tensor(ts, device=DEV)]
        assert_close(one_minus_beta, g[name + "_beta"], rtol=1e-4, atol=1e-5, what=f"{name} 1-beta")
