"""GPU, BASELINE.json full size (config 2: gridTD / VGG16, 64 images x 19 words = 1216 explanations, 224x224,
V = 10000, H = E = 512): the oracle cannot run this in seconds, so the checks are the size-independent properties the
rules offer — linearity of the relevance chain in its input relevance, independence of a request from the batch it
is computed in (bit-exact), relevance conservation from the decoder's output to the heat-map, word relevances
normalised to max |r| = 1 — plus equality of the CUDA-graph replay with the eager launches."""
import argparse
import os
import sys

import pytest
import torch

import synth
from conftest import ROOT

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def full():
    import bench
    from lrpx import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    args = argparse.Namespace(images=64, words=19, vocab=10000, chunk=128)
    model, ex, imgs, toks = bench.build_problem(args, torch.device(DEV), seed=0)
    imgs, toks = imgs.to(DEV), toks.to(DEV)
    eng, W = ex.engine(), ex._lrp_weights()
    B, T = 64, 19
    req_img = torch.arange(B, dtype=torch.int32, device=DEV).repeat_interleave(T)
    req_t = torch.arange(T, dtype=torch.int32, device=DEV).repeat(B)
    est = eng.forward(imgs)
    st = ex.explainer_forward(eng.features(est, "pixel"), toks)
    r_feat, r_words = ops.gridtd_decoder_lrp(st, W, req_img, req_t, toks[:, 1:].reshape(-1).to(torch.int32), tc_gemm=True)
    heat = eng.relevance(est, r_feat, req_img, chunk=128)
    torch.cuda.synchronize()
    return dict(ex=ex, eng=eng, est=est, imgs=imgs, toks=toks, req_img=req_img, req_t=req_t, r_feat=r_feat,
                r_words=r_words, heat=heat, B=B, T=T)


def test_outputs_finite_and_words_normalised(full):
    assert torch.isfinite(full["heat"]).all() and torch.isfinite(full["r_words"]).all()
    assert float(full["heat"].abs().max()) > 0
    # gridTDmodel.py:1129-1132: r_words / max|r_words| over the words 0..t of each request
    T = full["T"]
    mask = torch.arange(T, device=DEV)[None, :] <= full["req_t"][:, None]
    m = (full["r_words"].abs() * mask).amax(1)
    assert torch.allclose(m, torch.ones_like(m), atol=1e-6), (float(m.min()), float(m.max()))
    assert float((full["r_words"] * ~mask).abs().max()) == 0.0


def test_conservation_decoder_output_to_heatmap(full):
    """alpha=1/beta=0 without bias conserves relevance layer by layer (lrp_modules.py:81-84,134): the heat-map of a
    request sums to the relevance the decoder handed to the encoder output, up to bf16 rounding of the chain and the
    relevance dropped where z+ == 0.  Reported per request; bar: 1e-3 of sum |R| (measured: 7e-5)."""
    rin = full["heat"].double().flatten(1).sum(1)
    rout = full["r_feat"].double().flatten(1).sum(1)
    scale = full["r_feat"].double().flatten(1).abs().sum(1)
    rel = ((rin - rout).abs() / scale)
    print(f"conservation |sum R_in - sum R_out| / sum|R_out|: median {float(rel.median()):.3e} max {float(rel.max()):.3e}")
    assert float(rel.max()) < 1e-3


def test_chain_is_linear_in_the_relevance(full):
    eng, est = full["eng"], full["est"]
    q = slice(300, 428)                                     # one chunk's worth of requests
    r1, rimg = full["r_feat"][q], full["req_img"][q]
    g = torch.Generator(device=DEV).manual_seed(1)
    r2 = r1[torch.randperm(128, device=DEV, generator=g)] * 0.5
    h1 = eng.relevance(est, r1, rimg, chunk=128)
    h2 = eng.relevance(est, r2, rimg, chunk=128)
    h12 = eng.relevance(est, 2.0 * r1 - 3.0 * r2, rimg, chunk=128)
    ref = 2.0 * h1.double() - 3.0 * h2.double()
    err = float((h12.double() - ref).norm() / ref.norm())
    print(f"linearity rel L2 {err:.3e}")
    assert err < 5e-3                                        # bf16 storage of s between the 13 layers (measured 5e-4)
    assert torch.equal(h1, full["heat"][q])                  # same requests in another batch composition: bit-exact


def test_request_is_independent_of_its_batch(full):
    eng, est = full["eng"], full["est"]
    for q in (0, 517, 1215):
        alone = eng.relevance(est, full["r_feat"][q:q + 1], full["req_img"][q:q + 1], chunk=1)
        assert torch.equal(alone[0], full["heat"][q]), q


def test_graph_replay_equals_eager_at_full_size(full):
    from lrpx.pipeline import BatchExplainer
    pipe = BatchExplainer(full["ex"], chunk=128, use_graph=True)
    heat, words = pipe.explain(full["imgs"], full["toks"])
    heat, words = pipe.explain(full["imgs"], full["toks"])      # second call = pure replay
    torch.cuda.synchronize()
    assert torch.equal(heat, full["heat"])
    assert torch.equal(words, full["r_words"])


def test_config5_tuner_weights_and_step_at_full_size():
    """BASELINE config 5 (lrp_tune, batch 128 per GPU, V = 10000, H = E = 512, 224x224 images): the batched tuner
    kernel vs the oracle's restatement of get_lrp_weight_step (gridTDmodel.py:549-578) on the full-size logits,
    bit-exact argmax; then one whole train_lrp iteration (train.py:211-233) on the full model: finite losses,
    LRP weights in [0, 2], decoder parameters updated, encoder fixed."""
    import lrp_oracle as O
    from lrpx import ops, tune
    from models import gridTDmodel as G
    V, H, E, B, L = 10000, 512, 512, 128, 21
    g = torch.Generator().manual_seed(5)
    logits, h, c = torch.randn(B, V, generator=g), torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    W_fc = torch.randn(V, H, generator=g) * 0.05
    stop = synth.stop_mask(V)
    w_ctx, w_h, am = ops.fc_lrp_weights(logits.to(DEV), h.to(DEV), c.to(DEV), W_fc.to(DEV), stop.to(DEV))
    rc, rh = O.lrp_weight_step(logits, h, c, W_fc, stop)
    assert torch.equal(am.cpu().long(), logits.argmax(-1))
    assert float((w_ctx.cpu() - rc).abs().max()) < 1e-5 and float((w_h.cpu() - rh).abs().max()) < 1e-5
    assert float(w_ctx.min()) >= 0.0 and float(w_ctx.max()) <= 2.0
    # one full-size training iteration
    torch.manual_seed(0)
    model = G.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(1000, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(2000))
    model.to(DEV)
    st = tune.LrpTuneStep(model, synth.word_map(V), lr=1e-4, grad_clip=5.0)
    imgs = torch.randn(B, 3, 224, 224, generator=g).to(DEV)
    caps = torch.randint(1, V - 4, (B, L), generator=g)
    caps[:, 0] = V - 2
    enc_before = model.img_encoder.encoder[0].weight.detach().clone()
    fc_before = model.fc.weight.detach().clone()
    loss, ls, ll = st.step(imgs, caps.to(DEV), torch.full((B,), L))
    assert torch.isfinite(loss) and torch.isfinite(ls) and torch.isfinite(ll)
    assert torch.equal(enc_before, model.img_encoder.encoder[0].weight.detach())
    assert not torch.equal(fc_before, model.fc.weight.detach())
    # the fixed CNN ran on the tcgen05 forward (LrpTuneStep's default tc_encoder='fp32': bf16x3); the same batch through
    # the library convolutions gives the same losses to ~1e-4 (features within 1.5e-4 of fp32)
    assert model._tc_encoder is not None
    with torch.no_grad():
        l_tc = [float(v) for v in st.losses(imgs, caps.to(DEV), torch.full((B,), L))]
        model.use_tc_encoder(None)
        l_lib = [float(v) for v in st.losses(imgs, caps.to(DEV), torch.full((B,), L))]
    print("lrp_tune losses, tcgen05 encoder vs library encoder:", l_tc, l_lib)
    assert all(abs(a - b) <= 2e-3 * abs(b) for a, b in zip(l_tc, l_lib))


def test_adaptive_decoder_full_size_properties(tmp_path):
    """ExplainAdaptiveAttention at the bench size (64 images x 19 words, V = 10000, H = E = 512, P = 196, C = 512), where
    the oracle would need minutes: finite outputs, word relevances normalised to max |r| = 1, a request's result
    independent of the launch it is computed in (bit-exact), bf16x3 tensor-core GEMMs within 1e-4 (scale-relative) of the
    fp32 CUDA-core GEMMs (measured 8e-6)."""
    from lrpx import ops
    from models import adaptiveattention as AA
    V, H, E, B, T, P, C = 10000, 512, 512, 64, 19, 196, 512
    model = AA.AdaptiveAttentionCaptioningModel(E, H, V, "vgg16")
    model.load_state_dict(synth.adaptive_decoder_state(901, V, H, E), strict=False)
    model.to(DEV).eval()
    ns = argparse.Namespace(embed_dim=E, hidden_dim=H, encoder="vgg16", height=224, width=224, save_path=str(tmp_path),
                            dataset="syn", weight="")
    ex = AA.ExplainAdaptiveAttention(ns, synth.word_map(V), model=model, precision="fp32")
    g = torch.Generator().manual_seed(902)
    feat = torch.rand(B, P, C, generator=g).to(DEV)
    toks = torch.randint(1, V - 4, (B, T + 1), generator=g).to(DEV)
    st, W = ex.explainer_forward(feat, toks), ex._lrp_weights()
    req_img = torch.arange(B, dtype=torch.int32, device=DEV).repeat_interleave(T)
    req_t = torch.arange(T, dtype=torch.int32, device=DEV).repeat(B)
    req_word = toks[:, 1:].reshape(-1).to(torch.int32)
    f32_feat, f32_words = ops.adaptive_decoder_lrp(st, W, req_img, req_t, req_word, tc_gemm=False)
    tc_feat, tc_words = ops.adaptive_decoder_lrp(st, W, req_img, req_t, req_word, tc_gemm=True)
    assert torch.isfinite(f32_feat).all() and torch.isfinite(tc_feat).all()
    mask = torch.arange(T, device=DEV)[None, :] <= req_t[:, None]
    m = (f32_words.abs() * mask).amax(1)
    assert torch.allclose(m, torch.ones_like(m), atol=1e-6) and float((f32_words * ~mask).abs().max()) == 0.0
    scale = f32_feat.abs().flatten(1).amax(1)
    err = ((tc_feat - f32_feat).abs().flatten(1).amax(1) / scale).max()
    print(f"adaptive full size: bf16x3 vs fp32 GEMMs max scale-relative error {float(err):.3e}")
    assert float(err) <= 1e-4
    for q in (0, 611, 1215):
        one = ops.adaptive_decoder_lrp(st, W, req_img[q:q + 1], req_t[q:q + 1], req_word[q:q + 1], tc_gemm=False)
        assert torch.equal(one[0][0], f32_feat[q]) and torch.equal(one[1][0], f32_words[q]), q
