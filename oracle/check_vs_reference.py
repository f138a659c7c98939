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
    return 0


if __name__ == "__main__":
    sys.exit(main())
