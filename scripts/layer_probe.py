"""GPU probe: per-layer time of the relevance-chain launches under the experiment switches of conv_tc.cu."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
import bench
from lrpx import tc

sd = synth.vgg_state(1)
eng = tc.TcVggEngine([sd[k] for k in sd if k.endswith("weight")], [sd[k] for k in sd if k.endswith("bias")], synth.VGG16_CFG, "cuda")
eng.forward(torch.randn(1, 3, 224, 224, device="cuda"))
chunk = int(os.environ.get("CHUNK", "128"))
for name, env in [("tap", {"LRPX_TC_SLAB": "0"}), ("tap skip_epi_io", {"LRPX_TC_SLAB": "0", "LRPX_TC_DEBUG": "1"}),
                  ("slab", {"LRPX_TC_SLAB": "1"}), ("slab skip_epi_io", {"LRPX_TC_SLAB": "1", "LRPX_TC_DEBUG": "1"})]:
    for k in ("LRPX_TC_DEBUG", "LRPX_TC_SLAB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    rows = bench.layer_table(eng, chunk, torch.device("cuda"), None)
    print(name, " ".join(f"L{r['layer']}:{r['ms']:.3f}" for r in rows), "sum", round(sum(r["ms"] for r in rows), 3))
