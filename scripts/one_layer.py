"""GPU probe: launch the relevance-chain kernel of selected VGG16 layers a few times (for ncu captures).
   LAYERS="1,3" CHUNK=128 REPS=3 [LRPX_TC_SLAB=0|1] python scripts/one_layer.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from lrpx import tc

sd = synth.vgg_state(1)
PREC = os.environ.get("PRECISION", "bf16")          # fp32: the general (MULX) kernels with hi|lo rows
eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")], synth.VGG16_CFG, "cuda",
                     precision=PREC)
st = eng.forward(torch.randn(1, 3, 224, 224, device="cuda"))
n = int(os.environ.get("CHUNK", "128"))
reps = int(os.environ.get("REPS", "3"))
rimg = torch.zeros(n, dtype=torch.int32, device="cuda")
for li in [int(v) for v in os.environ.get("LAYERS", "1").split(",")]:
    c = eng.convs[li]
    a = torch.randn(tc.pf_rows(n, c.h, c.w), c.cout * eng.rm, device="cuda").to(torch.bfloat16)
    if eng.general and li > 0:
        below = eng.convs[li - 1]
        oh, ow = (2 * c.h, 2 * c.w) if below.pool_after else (c.h, c.w)
        out = torch.empty(tc.pf_rows(n, oh, ow), c.cin * eng.rm, device="cuda", dtype=torch.bfloat16)
        fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout * eng.km, c.cin, 3,
                                tc.EPI_MULX_UNPOOL if below.pool_after else tc.EPI_MULX, out, gain=st.gain[li - 1],
                                row_img=rimg, pool_idx=st.idx[li - 1] if below.pool_after else None,
                                a_phys=c.cout * eng.rm if eng.split else 0, groups=1, split=int(eng.split))
    elif li == 0:
        out = torch.empty(n, 3, c.h, c.w, device="cuda")
        if c.w_rel3 is not None:
            fn = lambda: tc.tc_conv(a, c.w_rel3, n, c.h, c.w, c.cout, 24, 3, tc.EPI_INPUT3, out, row_img=rimg, x=st.x)
        else:
            fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, 16, 3, tc.EPI_INPUT, out, row_img=rimg, x=st.x)
    elif eng.convs[li - 1].pool_after:
        out = torch.empty(tc.pf_rows(n, 2 * c.h, 2 * c.w), c.cin, device="cuda", dtype=torch.bfloat16)
        fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, c.cin, 3, tc.EPI_MUL_UNPOOL, out, gain=st.gain[li - 1],
                                row_img=rimg, pool_idx=st.idx[li - 1])
    else:
        out = torch.empty(tc.pf_rows(n, c.h, c.w), c.cin, device="cuda", dtype=torch.bfloat16)
        fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, c.cin, 3, tc.EPI_MUL, out, gain=st.gain[li - 1], row_img=rimg)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"layer {li}: min {ts[0]:.4f} med {ts[len(ts) // 2]:.4f} max {ts[-1]:.4f} ms (chunk {n})")
    del a, out
