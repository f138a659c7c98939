"""GPU parity: the tensor-core path (tcgen05 / TMA, bf16 operands, fp32 accumulation) through the C ABI.

Tolerances (stated here as north_star asks for the bf16 path):
  * GEMM machinery (STORE_F32 epilogue) vs an fp32 torch convolution of the SAME bf16-rounded operands:
    rtol 1e-4 / atol 1e-4 * max|ref| — only the fp32 accumulation order differs.
  * whole encoder chain vs the fp32 oracle / reference fixtures: Spearman rank correlation >= 0.99 on the
    heat-map and relative L2 error <= 5e-2 (bf16 storage of activations, gains and relevances between layers).
  * max-pool argmax bytes: bit-exact vs torch on the same bf16 values.
"""
import pytest
import torch
import torch.nn.functional as F

import lrp_oracle as O
import synth
from conftest import assert_close, spearman

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(x):
    return x.to(torch.bfloat16).float()


def _pf_valid(pf, n, h, w, c):
    """PF rows (n*(h+1)*(w+1), c) -> NCHW of the real pixels."""
    return pf.view(n, h + 1, w + 1, c)[:, 1:, 1:, :].permute(0, 3, 1, 2).float()


def _pf_pads(pf, n, h, w, c):
    v = pf.view(n, h + 1, w + 1, c).float()
    return torch.cat([v[:, 0].reshape(-1), v[:, :, 0].reshape(-1)])


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("slab", ["1", "0"])
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 8, 8, 64, 64), (3, 10, 6, 128, 256), (2, 14, 14, 64, 512),
                                            (5, 28, 28, 256, 128), (1, 4, 4, 512, 32), (2, 56, 56, 128, 128),
                                            (1, 112, 112, 64, 64), (1, 224, 224, 64, 64), (1, 224, 224, 64, 128),
                                            # >= 3 tiles per persistent CTA (ring phases wrap, both issuer warps busy):
                                            (8, 112, 112, 128, 64), (2, 224, 224, 64, 64), (2, 224, 224, 64, 32),
                                            (3, 224, 224, 64, 128), (40, 28, 28, 512, 256)])
def test_gemm_forward_layout(n, h, w, cin, cout, slab, monkeypatch):
    """slab=1: row-shifted descriptor views into one A slab per channel block (single slab, two-box slab,
    three-slab, B-resident and 256-row tile configurations are all hit by these shapes); slab=0: one TMA tile
    per filter tap."""
    from lrpx import tc
    monkeypatch.setenv("LRPX_TC_SLAB", slab)
    g = torch.Generator().manual_seed(n * 1000 + h + cin + cout)
    x = _bf(torch.randn(n, cin, h, w, generator=g)).to(DEV)
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.1).to(DEV)
    a = tc.nchw_to_pf(x)
    assert float(_pf_pads(a, n, h, w, cin).abs().max()) == 0.0
    assert torch.equal(_pf_valid(a, n, h, w, cin), x)
    out = torch.full((tc.pf_rows(n, h, w), cout), float("nan"), device=DEV)
    tc.tc_conv(a, tc.weight_prep(wt, 0), n, h, w, cin, cout, 3, tc.EPI_STORE_F32, out)
    ref = F.conv2d(x, wt, None, 1, 1)
    got = _pf_valid(out, n, h, w, cout)
    assert_close(got, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()), what="tc conv forward")


@pytest.mark.parametrize("slab", ["1", "0"])
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 6, 10, 64, 128), (1, 14, 14, 512, 512)])
def test_gemm_relevance_layout(n, h, w, cin, cout, slab, monkeypatch):
    """mode-2 weights: acc = W+^T * s (the transposed convolution of utils.lrp_backward, utils.py:29)."""
    from lrpx import tc
    monkeypatch.setenv("LRPX_TC_SLAB", slab)
    g = torch.Generator().manual_seed(h * 31 + cout)
    s = _bf(torch.randn(n, cout, h, w, generator=g)).to(DEV)
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.1).to(DEV)
    out = torch.empty(tc.pf_rows(n, h, w), cin, device=DEV)
    tc.tc_conv(tc.nchw_to_pf(s), tc.weight_prep(wt, 2), n, h, w, cout, cin, 3, tc.EPI_STORE_F32, out)
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt.clamp(min=0), s, 1, 1)
    assert_close(_pf_valid(out, n, h, w, cin), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()), what="tc dgrad")


def test_1x1_and_many_tiles():
    from lrpx import tc
    g = torch.Generator().manual_seed(77)
    n, h, w, cin, cout = 40, 14, 14, 64, 64          # 40*225 rows = 71 M tiles > one wave is not needed; ragged tail
    x = _bf(torch.randn(n, cin, h, w, generator=g)).to(DEV)
    wt = _bf(torch.randn(cout, cin, 1, 1, generator=g)).to(DEV)
    out = torch.empty(tc.pf_rows(n, h, w), cout, device=DEV)
    tc.tc_conv(tc.nchw_to_pf(x), tc.weight_prep(wt, 0), n, h, w, cin, cout, 1, tc.EPI_STORE_F32, out)
    ref = F.conv2d(x, wt)
    assert_close(_pf_valid(out, n, h, w, cout), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()), what="tc 1x1")
    # enough tiles that every CTA of the persistent grid loops more than once (double-buffered accumulators)
    n = 400
    x = _bf(torch.randn(n, cin, h, w, generator=g)).to(DEV)
    out = torch.empty(tc.pf_rows(n, h, w), cout, device=DEV)
    tc.tc_conv(tc.nchw_to_pf(x), tc.weight_prep(wt, 0), n, h, w, cin, cout, 1, tc.EPI_STORE_F32, out)
    assert_close(_pf_valid(out, n, h, w, cout), F.conv2d(x, wt), rtol=1e-4, atol=1e-4 * float(ref.abs().max()),
                 what="tc 1x1 many tiles")


def test_first_layer_and_maxpool():
    from lrpx import tc, _lib
    import ctypes as C
    g = torch.Generator().manual_seed(5)
    n, h, w, cout = 2, 12, 8, 64
    x = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(cout, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(cout, generator=g) * 0.1
    act = torch.empty(tc.pf_rows(n, h, w), cout, device=DEV, dtype=torch.bfloat16)
    gain = torch.empty_like(act)
    xd, wd, bd = x.to(DEV), wt.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().lrpx_tc_first_fwd(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), act.data_ptr(), gain.data_ptr(),
                                            n, h, w, cout, torch.cuda.current_stream().cuda_stream), "first_fwd")
    z = F.conv2d(x, wt, b, 1, 1)
    zp = F.conv2d(x.clamp(min=0), wt.clamp(min=0), None, 1, 1) + F.conv2d(x.clamp(max=0), wt.clamp(max=0), None, 1, 1)
    a_ref = z.clamp(min=0)
    assert_close(_pf_valid(act, n, h, w, cout), a_ref, rtol=1e-2, atol=1e-2, what="first layer act (bf16)")
    g_ref = O.safe_divide(a_ref, zp)
    assert_close(_pf_valid(gain, n, h, w, cout), g_ref, rtol=2e-2, atol=2e-2, what="first layer gain (bf16)")
    assert float(_pf_pads(act, n, h, w, cout).abs().max()) == 0.0
    # 2x2 max-pool with argmax + gain gather, bit-exact vs torch on the same bf16 values
    pooled, idx, gpool = tc.maxpool2(act, gain, n, h, w, cout)
    av, gv = _pf_valid(act, n, h, w, cout), _pf_valid(gain, n, h, w, cout)
    pref, iref = F.max_pool2d(av, 2, 2, return_indices=True)
    assert torch.equal(_pf_valid(pooled, n, h // 2, w // 2, cout), pref)
    yy, xx = iref // w, iref % w
    k_ref = ((yy % 2) * 2 + (xx % 2)).to(torch.uint8)
    assert torch.equal(_pf_valid(idx, n, h // 2, w // 2, cout).to(torch.uint8), k_ref), "argmax must be bit-exact"
    g_at = gv.flatten(2).gather(2, iref.flatten(2)).view_as(pref)
    assert torch.equal(_pf_valid(gpool, n, h // 2, w // 2, cout), g_at)


def _engine_vs_oracle(cfg, seed, n, size, chunk):
    from lrpx import tc
    sd = synth.vgg_state(seed, cfg)
    ws = [sd[k] for k in sd if k.endswith("weight")]
    bs = [sd[k] for k in sd if k.endswith("bias")]
    eng = tc.TcVggEngine(ws, bs, cfg, DEV)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, 3, size, size, generator=g)
    st = eng.forward(x.to(DEV))
    layers = O.vgg_layers_from_state(sd, cfg)
    feats = O.sequential_forward(layers, x)[-1]
    got_feats = eng.features(st, "nchw")
    rel_err = float((got_feats.cpu() - feats).norm() / feats.norm())
    print(f"forward rel L2 err {rel_err:.3e}")
    assert rel_err < 3e-2
    # two requests per image, with different relevance targets
    Q = 2 * n
    row_img = torch.arange(Q, dtype=torch.int32) % n
    C, fh, fw = feats.shape[1:]
    tgt = torch.randn(Q, C, fh, fw, generator=g) * feats[row_img.long()]       # relevance ~ a (.) something
    r_pix = tgt.flatten(2).transpose(1, 2).contiguous()                       # (Q, P, C)
    heat = eng.relevance(st, r_pix.to(DEV), row_img.to(DEV), chunk=chunk)
    layers64 = [tuple(v.double() if torch.is_tensor(v) else v for v in l) for l in layers]
    ref = O.sequential_lrp(layers64, x[row_img.long()].double(), tgt.double())
    for q in range(Q):
        a, b = heat[q].cpu().double(), ref[q]
        l2 = float((a - b).norm() / b.norm())
        sp = spearman(a, b)
        print(f"request {q}: rel L2 {l2:.3e} spearman {sp:.5f} sumR {float(a.sum()):.5g} vs {float(b.sum()):.5g}")
        assert sp >= 0.99, sp
        assert l2 <= 5e-2, l2
    return heat


def test_engine_small_vs_oracle():
    _engine_vs_oracle([64, 64, "M", 128, 128, "M", 256], seed=3, n=3, size=16, chunk=4)


def test_engine_pool_chain_vs_oracle():
    _engine_vs_oracle([64, "M", 64, "M", 128, "M", 128, 128], seed=4, n=2, size=32, chunk=64)


def test_engine_vgg16_224_vs_reference_fixture(golden):
    """BASELINE size; the fixture is the reference's own compute_lrp output (fp32, CPU)."""
    from lrpx import tc
    g = golden("vgg16_224")
    seed = int(g["seed"])
    sd = synth.vgg_state(seed)
    eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")],
                         synth.VGG16_CFG, DEV)
    gen = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(1, 3, 224, 224, generator=gen)
    tgt = torch.randn(1, 512, 14, 14, generator=gen) * 1e-3
    st = eng.forward(x.to(DEV))
    heat = eng.relevance(st, tgt.flatten(2).transpose(1, 2).contiguous().to(DEV))
    a, b = heat[0].cpu().double(), g["rel"][0].double()
    l2 = float((a - b).norm() / b.norm())
    sp = spearman(a, b)
    print(f"vgg16 224: rel L2 {l2:.3e} spearman {sp:.5f} sumR {float(a.sum()):.6g} vs {float(b.sum()):.6g}")
    assert sp >= 0.99 and l2 <= 5e-2
    # the same request 6 times in one chunk: every persistent CTA now walks several tiles per layer (accumulator
    # double-buffering, ring wrap-around, both MMA issuer warps); each copy must reproduce the single-request result
    # bit for bit (same arithmetic, different tile -> CTA assignment)
    Q = 6
    r6 = tgt.flatten(2).transpose(1, 2).contiguous().repeat(Q, 1, 1).to(DEV)
    heat6 = eng.relevance(st, r6, torch.zeros(Q, dtype=torch.int32, device=DEV), chunk=Q)
    for q in range(Q):
        assert torch.equal(heat6[q], heat[0]), f"request {q} of the replicated chunk differs from the single request"


def test_relevance_groups_and_chunks_do_not_change_results():
    """TcVggEngine.relevance over many requests: the stage-1 grouping (GROUP), the chunk size and the chunk-wise
    callback must not change a single bit of any request's heat-map, and the callback must cover every request once."""
    from lrpx import tc
    cfg = [64, 64, "M", 128, "M", 128, 128]
    sd = synth.vgg_state(7, cfg)
    eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")], cfg, DEV)
    g = torch.Generator().manual_seed(8)
    n, Q = 3, 11
    st = eng.forward(torch.randn(n, 3, 32, 32, generator=g).to(DEV))
    row_img = (torch.arange(Q, dtype=torch.int32) % n).to(DEV)
    r = torch.randn(Q, 64, 128, generator=g).to(DEV)
    ref = eng.relevance(st, r, row_img, chunk=Q)                 # one chunk, no stage 1
    for group, chunk in ((4, 2), (6, 3), (2048, 4), (5, 5)):
        eng.GROUP = group
        seen = []
        got = eng.relevance(st, r, row_img, chunk=chunk, on_chunk=lambda q0, q1: seen.append((q0, q1)))
        assert torch.equal(got, ref), (group, chunk)
        assert seen[0][0] == 0 and seen[-1][1] == Q and all(a[1] == b[0] for a, b in zip(seen, seen[1:])), seen
    eng.GROUP = 2048


@pytest.mark.parametrize("n,h,w", [(1, 8, 8), (3, 12, 20), (2, 56, 56), (5, 224, 224), (2, 37, 201), (1, 300, 251),
                                   (3, 9, 230)])
def test_first_layer_folded_columns_equals_plain_form(n, h, w, monkeypatch):
    """LRPX_TC_EPI_INPUT3 (filter columns folded into N, column shift in the epilogue, 126-row M halves) vs
    LRPX_TC_EPI_INPUT (one MMA chain per filter tap) on the same operands: same bf16 products, fp32 accumulation in
    another order -> equal within 1e-5 of the largest element; every pixel of every request is written."""
    from lrpx import tc
    g = torch.Generator().manual_seed(h * 7 + w)
    cout = 64
    wt = torch.randn(cout, 3, 3, 3, generator=g) * 0.2
    sd = {"0.weight": wt, "0.bias": torch.zeros(cout)}
    eng = tc.TcVggEngine([wt], [torch.zeros(cout)], [cout], DEV)
    c0 = eng.convs[0]
    assert c0.w_rel3 is not None
    x = torch.randn(n, 3, h, w, generator=g).to(DEV)
    s = _bf(torch.randn(n, cout, h, w, generator=g)).to(DEV)
    a = tc.nchw_to_pf(s)
    rimg = torch.arange(n, dtype=torch.int32, device=DEV)
    out16 = torch.full((n, 3, h, w), float("nan"), device=DEV)
    out24 = torch.full((n, 3, h, w), float("nan"), device=DEV)
    tc.tc_conv(a, c0.w_rel, n, h, w, cout, 16, 3, tc.EPI_INPUT, out16, row_img=rimg, x=x)
    tc.tc_conv(a, c0.w_rel3, n, h, w, cout, 24, 3, tc.EPI_INPUT3, out24, row_img=rimg, x=x)
    assert torch.isfinite(out24).all()
    # switches of the same kernel: two accumulator buffers instead of four (LRPX_TC_NBUF=2), one ring stage per slab instead
    # of one per tile (LRPX_TC_AGROUP=0; wide images) and the row walk (tiles one image row apart, two of three slabs reused;
    # LRPX_TC_WALK=1, eligible for 201 <= w + 1 <= 252) issue the same MMAs in the same order -> bit-identical
    for key, val in (("LRPX_TC_NBUF", "2"), ("LRPX_TC_AGROUP", "0"), ("LRPX_TC_WALK", "1")):
        monkeypatch.setenv(key, val)
        out24s = torch.full((n, 3, h, w), float("nan"), device=DEV)
        tc.tc_conv(a, c0.w_rel3, n, h, w, cout, 24, 3, tc.EPI_INPUT3, out24s, row_img=rimg, x=x)
        monkeypatch.delenv(key)
        assert torch.equal(out24, out24s), key
    scale = float(out16.abs().max())
    assert float((out24 - out16).abs().max()) <= 1e-5 * scale
    # and against the rule itself: R = x+ (W+^T * s) + x- (W-^T * s)   (lrp_modules.py:81-84, utils.py:26-30)
    wb = _bf(wt).to(DEV)
    cp = torch.nn.grad.conv2d_input((n, 3, h, w), wb.clamp(min=0), s, 1, 1)
    cn = torch.nn.grad.conv2d_input((n, 3, h, w), wb.clamp(max=0), s, 1, 1)
    ref = x.clamp(min=0) * cp + x.clamp(max=0) * cn
    assert_close(out24, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()), what="folded first layer vs rule")


def test_folded_filter_columns_equal_the_plain_contraction():
    """LRPX_TC_FOLD_COLUMNS (the 64-column MUL layer with its three filter columns folded into N, off by default in the
    engine): same result as the nine-tap contraction up to the fp32 summation order (bf16 output: 1 ulp)."""
    from lrpx import tc
    g = torch.Generator().manual_seed(5)
    n, h, w, cout, cin = 3, 20, 28, 64, 64
    wgt = torch.randn(cout, cin, 3, 3, generator=g).to(DEV)
    w_rel = tc.weight_prep(wgt, 2)                                                   # (cin, 9*cout)
    w_fold = w_rel.reshape(cin, 3, 3, cout).permute(2, 0, 1, 3).reshape(3 * cin, 3 * cout).contiguous()
    s = tc.nchw_to_pf(torch.randn(n, cout, h, w, generator=g).to(DEV))
    gain = tc.nchw_to_pf(torch.rand(2, cin, h, w, generator=g).to(DEV))
    rimg = torch.tensor([1, 0, 1], dtype=torch.int32, device=DEV)
    rows = tc.pf_rows(n, h, w)
    a = torch.empty(rows, cin, device=DEV, dtype=torch.bfloat16)
    b = torch.full((rows, cin), 7.0, device=DEV, dtype=torch.bfloat16)
    tc.tc_conv(s, w_rel, n, h, w, cout, cin, 3, tc.EPI_MUL, a, gain=gain, row_img=rimg)
    tc.tc_conv(s, w_fold, n, h, w, cout, cin, 3, tc.EPI_MUL, b, gain=gain, row_img=rimg, fwd_flags=8)
    assert torch.isfinite(b.float()).all()
    err = (a.float() - b.float()).abs()
    assert float(err.max()) <= 2.0 ** -7 * float(a.float().abs().max())
    assert float((err > 0).float().mean()) < 0.2                                      # most elements round identically
    pad = tc.pf_to_dense(b, n, h, w, cin)                                             # and the padding rows stay zero
    assert float(b.float().abs().sum()) == pytest.approx(float(pad.abs().sum()), rel=1e-3)


@pytest.mark.parametrize("n,h,w,cout,cin,unpool", [(6, 56, 56, 128, 128, False), (6, 28, 28, 256, 128, True),
                                                   (3, 112, 112, 64, 64, False), (3, 56, 56, 128, 64, True),
                                                   (40, 14, 14, 512, 256, False)])
def test_chain_layer_switches_do_not_change_results(n, h, w, cout, cin, unpool, monkeypatch):
    """One relevance-chain layer (EPI_MUL / EPI_MUL_UNPOOL, bf16) under the kernel's scheduling switches — no
    `cta_group::2` pair (LRPX_TC_PAIR=0), no clusters (LRPX_TC_CLUSTER=0), plain stores instead of TMA tile stores
    (LRPX_TC_TMASTORE=0), four accumulator buffers where they fit (LRPX_TC_NBUF=4), 128-row tiles (LRPX_TC_MH=1): every
    variant issues the same MMAs in the same K order, so the bf16 outputs are bit-identical."""
    from lrpx import tc
    g = torch.Generator().manual_seed(n * 100 + h + cout)
    s = _bf(torch.randn(n, cout, h, w, generator=g)).to(DEV)
    wt = torch.randn(cout, cin, 3, 3, generator=g).to(DEV) * 0.1
    a = tc.nchw_to_pf(s)
    w_rel = tc.weight_prep(wt, 2)
    oh, ow = (2 * h, 2 * w) if unpool else (h, w)
    gain = torch.rand(tc.pf_rows(n, h, w), cin, generator=g).to(torch.bfloat16).to(DEV)
    idx = torch.randint(0, 4, (tc.pf_rows(n, h, w), cin), generator=g, dtype=torch.uint8).to(DEV) if unpool else None
    rimg = torch.arange(n, dtype=torch.int32, device=DEV)

    def run():
        out = torch.full((tc.pf_rows(n, oh, ow), cin), float("nan"), device=DEV, dtype=torch.bfloat16)
        tc.tc_conv(a, w_rel, n, h, w, cout, cin, 3, tc.EPI_MUL_UNPOOL if unpool else tc.EPI_MUL, out, gain=gain,
                   row_img=rimg, pool_idx=idx)
        torch.cuda.synchronize()
        return out

    ref = run()
    assert torch.isfinite(_pf_valid(ref, n, oh, ow, cin)).all()
    for key, val in (("LRPX_TC_PAIR", "0"), ("LRPX_TC_CLUSTER", "0"), ("LRPX_TC_TMASTORE", "0"), ("LRPX_TC_NBUF", "4"),
                     ("LRPX_TC_MH", "1")):
        monkeypatch.setenv(key, val)
        got = run()
        monkeypatch.delenv(key)
        assert torch.equal(_pf_valid(got, n, oh, ow, cin), _pf_valid(ref, n, oh, ow, cin)), key
