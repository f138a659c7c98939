"""GPU debug probe: TcResNetEngine stage by stage against the oracle's rules (relevance at every block input)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lrp-imagecaptioning-pytorch_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch, torch.nn.functional as F
import synth, lrp_oracle as O
from lrpx import tc, tc_resnet
from models import resnet
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
layers, seed, size = (1, 1, 1, 1), 77, int(os.environ.get("SIZE", "64"))
sd = synth.resnet_state(seed, layers)
net = resnet.ResNet(resnet.Bottleneck, list(layers)); net.load_state_dict(sd); net = net.to(DEV).eval()
eng = tc_resnet.TcResNetEngine(net, DEV)
g = torch.Generator().manual_seed(78)
x = torch.randn(1, 3, size, size, generator=g)
st = eng.forward(x.to(DEV))
fh, fw = st.feat_hw
tgt = torch.randn(1, 2048, fh, fw, generator=g) * 1e-3
if os.environ.get("TGT", "feat") == "feat":       # relevance as a decoder hands it over: proportional to the feature
    tgt = tgt * eng.features(st, "nchw").cpu()
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
# ---- oracle with intermediates (copy of O.resnet_lrp's loop)
blocks = O.resnet_block_names(sd64)
xd = x.double()
c1 = F.conv2d(xd, sd64["conv1.weight"], None, 2, 3)
b1 = O._bn_apply(c1, sd64, "bn1").clamp(min=0)
mp = F.max_pool2d(b1, 3, 2, 1)
saved, cur = [], mp
for name in blocks:
    stride = 1 if (name.endswith(".0") is False or name.startswith("layer1")) else 2
    rec = {"in": cur, "stride": stride}
    o1 = F.conv2d(cur, sd64[name + ".conv1.weight"]); rec["c1"] = o1
    a1 = O._bn_apply(o1, sd64, name + ".bn1").clamp(min=0); rec["a1"] = a1
    o2 = F.conv2d(a1, sd64[name + ".conv2.weight"], None, stride, 1); rec["c2"] = o2
    a2 = O._bn_apply(o2, sd64, name + ".bn2").clamp(min=0); rec["a2"] = a2
    o3 = F.conv2d(a2, sd64[name + ".conv3.weight"]); rec["c3"] = o3
    y3 = O._bn_apply(o3, sd64, name + ".bn3"); rec["y3"] = y3
    if name + ".downsample.0.weight" in sd64:
        od = F.conv2d(cur, sd64[name + ".downsample.0.weight"], None, stride); rec["cd"] = od
        idn = O._bn_apply(od, sd64, name + ".downsample.1")
    else:
        idn = cur
    rec["idn"] = idn
    cur = (y3 + idn).clamp(min=0)
    saved.append(rec)
r = tgt.double()
rin = {}
for k in range(len(blocks) - 1, -1, -1):
    name, rec = blocks[k], saved[k]
    r_y3, r_idn = O.add_rule(rec["y3"], rec["idn"], r)
    rr = O._bn_rule(rec["c3"], r_y3, sd64, name + ".bn3")
    rr = O.conv_alpha_beta(rec["a2"], sd64[name + ".conv3.weight"], None, rr); rec["R_a2"] = rr
    rr = O._bn_rule(rec["c2"], rr, sd64, name + ".bn2")
    rr = O.conv_alpha_beta(rec["a1"], sd64[name + ".conv2.weight"], None, rr, rec["stride"], 1); rec["R_a1"] = rr
    rr = O._bn_rule(rec["c1"], rr, sd64, name + ".bn1")
    rr = O.conv_alpha_beta(rec["in"], sd64[name + ".conv1.weight"], None, rr); rec["R_main"] = rr
    if "cd" in rec:
        rd = O._bn_rule(rec["cd"], r_idn, sd64, name + ".downsample.1")
        rd = O.conv_alpha_beta(rec["in"], sd64[name + ".downsample.0.weight"], None, rd, rec["stride"], 0)
    else:
        rd = r_idn
    rec["R_id"] = rd
    r = rr + rd
    rin[k] = r
def rep(tag, a, b):
    a, b = a.cpu().double(), b.cpu().double()
    print(f"{tag}: rel L2 {float((a - b).norm() / b.norm()):.3e}  max|b| {float(b.abs().max()):.3e} sum {float(a.sum()):.5g} vs {float(b.sum()):.5g}")
def pfv(pf, h, w, c):
    return pf.view(1, h + 1, w + 1, c)[:, 1:, 1:, :].permute(0, 3, 1, 2).float()
rep("features", eng.features(st, "nchw"), cur)
r_pix = tgt.flatten(2).transpose(1, 2).contiguous().to(DEV)
rimg = torch.zeros(1, dtype=torch.int32, device=DEV)
for k in range(len(blocks) - 1, -1, -1):
    Rk = eng._blocks_relevance(st, r_pix, rimg, 1, stop_at=k)
    h, w = st.blocks[k]["hw"]
    rep(f"R at input of block {k} ({blocks[k]})", pfv(Rk, h, w, eng.blocks[k]["c1"].cin), rin[k])
heat = eng.relevance(st, r_pix)
ref = O.resnet_lrp(sd64, xd, tgt.double())
rep("heat", heat, ref)
# stem alone on the oracle's R0
R0 = tc.nchw_to_pf(rin[0].float().to(DEV))
h, w = st.stem["hw"]
import ctypes as C
from lrpx._lib import lib, check
P_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
S_ = C.c_void_p(torch.cuda.current_stream().cuda_stream)
A = torch.empty(tc.pf_rows(1, h, w), 64, device=DEV, dtype=torch.bfloat16)
check(lib().lrpx_tc_unpool3s2_bf16(P_(R0), P_(st.stem["idx"]), P_(st.stem["G"]), P_(rimg), P_(A), 1, h, w, 64, S_), "unpool")
Pm = torch.empty(tc.pf_rows(1, h, w), 320, device=DEV)
tc.tc_conv(A, eng.stem_w_rel, 1, h, w, 64, 320, 1, tc.EPI_STORE_F32, Pm)
out = torch.empty(1, 3, size, size, device=DEV)
check(lib().lrpx_tc_stem_col2im_f32(P_(Pm), 320, P_(st.x), P_(rimg), P_(out), 1, size, size, 0, S_), "col2im")
rep("stem alone (oracle R0 in)", out, ref)
rm, _ = O.maxpool_rule(b1, rin[0], 3, 2, 1)
rb = O._bn_rule(c1, rm, sd64, "bn1")
zp = O._conv_signed_net(xd, sd64["conv1.weight"], None, True, 2, 3, 1, 1, True)[0]
rep("stem operand A (= R_bn1 ratio / z+)", pfv(A, h, w, 64), O.safe_divide(rb, zp))

# ---- last block, tensor by tensor
print("---- last block operands")
k = len(blocks) - 1
blk, gk, rec, name = eng.blocks[k], st.blocks[k], saved[k], blocks[k]
c1_, c2_, c3_, dn_, s_ = blk["c1"], blk["c2"], blk["c3"], blk["down"], blk["stride"]
(h, w), (hc, wc) = gk["hw"], gk["hwc"]
zp = lambda a, wname, stride=1, pad=0: O._conv_signed_net(a, sd64[wname], None, True, stride, pad, 1, 1, True)[0]
r_top = tgt.double()
r_y3, r_idn = O.add_rule(rec["y3"], rec["idn"], r_top)
zp3 = zp(rec["a2"], name + ".conv3.weight")
A3_ref = O.safe_divide(O._bn_rule(rec["c3"], r_y3, sd64, name + ".bn3"), zp3)
bf = lambda rows, c: torch.empty(rows, c, device=DEV, dtype=torch.bfloat16)
C_ = st.feat_c
A3, S = bf(tc.pf_rows(1, fh, fw), C_), bf(tc.pf_rows(1, fh, fw), C_)
for rz, dst in ((gk["G3"], A3), (gk["Gs"], S)):
    check(lib().lrpx_tc_scale_rows(P_(r_pix), P_(rz), P_(rimg), P_(dst), 1, fh, fw, C_, S_), "scale_rows")
rep("A3 = R_c3 / z+3", pfv(A3, fh, fw, C_), A3_ref)
G3_ref = O.safe_divide(O._bn_rule(rec["c3"], O.add_rule(rec["y3"], rec["idn"], torch.ones_like(r_top))[0], sd64, name + ".bn3"), zp3)
rep("G3", pfv(gk["G3"], fh, fw, C_), G3_ref)
if "cd" in rec:
    zpd = zp(rec["in"], name + ".downsample.0.weight", rec["stride"], 0)
    S_ref = O.safe_divide(O._bn_rule(rec["cd"], r_idn, sd64, name + ".downsample.1"), zpd)
    rep("S = R_cd / z+d", pfv(S, fh, fw, C_), S_ref)
zp2 = zp(rec["a1"], name + ".conv2.weight", rec["stride"], 1)
A2_ref_c = O.safe_divide(O._bn_rule(rec["c2"], rec["R_a2"], sd64, name + ".bn2"), zp2)
A2 = bf(tc.pf_rows(1, h, w), c2_.cout)
tc.tc_conv(A3, c3_.w_rel, 1, hc, wc, c3_.cout, c3_.cin, 1, tc.EPI_MULX_UNPOOL if s_ == 2 else tc.EPI_MULX, A2, gain=gk["G2"], row_img=rimg)
A2v = pfv(A2, h, w, c2_.cout)
if s_ == 2:
    rep("A2 (even pixels)", A2v[:, :, ::2, ::2], A2_ref_c)
    print("A2 off-grid max", float(A2v[:, :, 1::2, :].abs().max()), float(A2v[:, :, :, 1::2].abs().max()))
else:
    rep("A2", A2v, A2_ref_c)
rep("G2", pfv(gk["G2"], hc, wc, c2_.cout), O.safe_divide(O._bn_rule(rec["c2"], rec["a2"], sd64, name + ".bn2"), zp2))
zp1 = zp(rec["in"], name + ".conv1.weight")
A1_ref = O.safe_divide(O._bn_rule(rec["c1"], rec["R_a1"], sd64, name + ".bn1"), zp1)
A1 = bf(tc.pf_rows(1, h, w), c1_.cout)
tc.tc_conv(A2, c2_.w_rel, 1, h, w, c2_.cout, c2_.cin, 3, tc.EPI_MULX, A1, gain=gk["G1"], row_img=rimg)
rep("A1", pfv(A1, h, w, c1_.cout), A1_ref)
if dn_ is not None:
    add = bf(tc.pf_rows(1, h, w), dn_.cin)
    tc.tc_conv(S, dn_.w_rel, 1, hc, wc, dn_.cout, dn_.cin, 1, tc.EPI_MULX_UNPOOL if s_ == 2 else tc.EPI_MULX, add, gain=gk["xs"], row_img=rimg)
    rep("add = R_id", pfv(add, h, w, dn_.cin), rec["R_id"])
R0 = bf(tc.pf_rows(1, h, w), c1_.cin)
tc.tc_conv(A1, c1_.w_rel, 1, h, w, c1_.cout, c1_.cin, 1, tc.EPI_MULX, R0, gain=gk["x"], groups=1, row_img=rimg)
rep("R_main", pfv(R0, h, w, c1_.cin), rec["R_main"])
