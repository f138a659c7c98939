"""Seeded synthetic weights / inputs shared by the tests, the golden generator and bench.py.

Weights are NOT committed: every fixture stores the seed, and both sides (the reference
when ``oracle/make_golden.py`` ran, and the CUDA path / oracle in the tests) rebuild the
same tensors from a ``torch.Generator`` on the CPU.  Key names follow the reference's
``state_dict`` layout (models/gridTDmodel.py:106-131, models/aoamodel.py:111-139,
models/vgg.py:62-83, models/resnet.py:95-239) so the dicts load into the reference
modules with ``load_state_dict``.
"""
import math
from typing import Dict, List, Sequence

import torch

VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def vgg_state(seed: int, cfg: Sequence = VGG16_CFG, in_ch: int = 3, prefix: str = "", bias_scale: float = 0.05):
    """features.* of a VGG-style Sequential (conv3x3 pad1 + ReLU, 'M' = maxpool 2x2)."""
    g = _gen(seed)
    sd, idx, c = {}, 0, in_ch
    for v in cfg:
        if v == "M":
            idx += 1
            continue
        std = math.sqrt(2.0 / (v * 9))          # kaiming_normal_, fan_out (vgg.py:48-52)
        sd[f"{prefix}{idx}.weight"] = torch.randn(v, c, 3, 3, generator=g) * std
        sd[f"{prefix}{idx}.bias"] = torch.randn(v, generator=g) * bias_scale
        c = v
        idx += 2
    return sd


def _linear(g, sd, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(g, (out_f, in_f), b)
    if bias:
        sd[name + ".bias"] = _uniform(g, (out_f,), b)


def _lstm(g, sd, name, in_f, hid):
    b = 1.0 / math.sqrt(hid)
    sd[name + ".weight_ih"] = _uniform(g, (4 * hid, in_f), b)
    sd[name + ".weight_hh"] = _uniform(g, (4 * hid, hid), b)
    sd[name + ".bias_ih"] = _uniform(g, (4 * hid,), b)
    sd[name + ".bias_hh"] = _uniform(g, (4 * hid,), b)


def gridtd_decoder_state(seed: int, V: int, H: int, E: int, C: int = 512, n_pixel: int = 196):
    """Everything in GridTDModel except img_encoder.* (gridTDmodel.py:106-131)."""
    g = _gen(seed)
    sd = {}
    b = 1.0 / math.sqrt(C)
    sd["img_projector.weight"] = _uniform(g, (H, C, 1, 1), b)
    sd["img_projector.bias"] = _uniform(g, (H,), b)
    _linear(g, sd, "global_img_feature_proj", E, C)
    _lstm(g, sd, "LanguageLSTM", 2 * H, H)
    _lstm(g, sd, "AdaLSTM.lstm_cell", 2 * E + H, H)
    _linear(g, sd, "AdaLSTM.x_gate", H, 2 * E + H)
    _linear(g, sd, "AdaLSTM.h_gate", H, H)
    _linear(g, sd, "AdaAttention.W_v_proj", n_pixel, H)
    _linear(g, sd, "AdaAttention.W_s_proj", n_pixel, H)
    _linear(g, sd, "AdaAttention.W_g_proj", n_pixel, H, bias=False)
    _linear(g, sd, "AdaAttention.w_h", 1, n_pixel, bias=False)
    sd["embedding.weight"] = torch.randn(V, E, generator=g)
    _linear(g, sd, "fc", V, H)
    return sd


def adaptive_decoder_state(seed: int, V: int, H: int, E: int, C: int = 512, n_pixel: int = 196):
    """Everything in AdaptiveAttentionCaptioningModel except img_encoder.* (adaptiveattention.py:103-121);
    the explainer's saved x_t is sized embed_dim + hidden_dim (:636), so E must equal H."""
    g = _gen(seed)
    sd = {}
    b = 1.0 / math.sqrt(C)
    sd["img_projector.weight"] = _uniform(g, (H, C, 1, 1), b)
    sd["img_projector.bias"] = _uniform(g, (H,), b)
    _linear(g, sd, "global_img_feature_proj", E, C)
    _lstm(g, sd, "AdaLSTM.lstm_cell", 2 * E, H)
    _linear(g, sd, "AdaLSTM.x_gate", H, 2 * E)
    _linear(g, sd, "AdaLSTM.h_gate", H, H)
    _linear(g, sd, "AdaAttention.W_v_proj", n_pixel, H)
    _linear(g, sd, "AdaAttention.W_s_proj", n_pixel, H)
    _linear(g, sd, "AdaAttention.W_g_proj", n_pixel, H, bias=False)
    _linear(g, sd, "AdaAttention.w_h", 1, n_pixel, bias=False)
    sd["embedding.weight"] = torch.randn(V, E, generator=g)
    _linear(g, sd, "fc", V, H)
    return sd


def aoa_decoder_state(seed: int, V: int, H: int, E: int, C: int = 512):
    """Everything in AOAModel except img_encoder.* (aoamodel.py:111-139)."""
    g = _gen(seed)
    sd = {}
    b = 1.0 / math.sqrt(C)
    sd["img_projector.weight"] = _uniform(g, (H, C, 1, 1), b)
    sd["img_projector.bias"] = _uniform(g, (H,), b)
    sd["embedding.weight"] = torch.randn(V, E, generator=g)
    _lstm(g, sd, "LanguageLSTM", H + E, H)
    _linear(g, sd, "decoder_k_proj", H, H)
    _linear(g, sd, "decoder_v_proj", H, H)
    _linear(g, sd, "decoder_multihead_attention.q_proj", H, H)
    _linear(g, sd, "decoder_aoa_linear_gate", H, H)
    _linear(g, sd, "decoder_aoa_linear", H, H)
    _linear(g, sd, "fc", V, H)
    return sd


def gridtd_bu_state(seed: int, V: int, H: int, E: int, n_region: int = 36, C: int = 2048):
    """GridTDModelBU (gridTDmodel.py:1868-1886): Linear projector over C-d regions, global projection hidden -> embed."""
    sd = gridtd_decoder_state(seed, V, H, E, C=C, n_pixel=n_region)
    sd["img_projector.weight"] = sd["img_projector.weight"].reshape(H, C).clone()
    g = _gen(seed + 100003)
    for k in ("global_img_feature_proj.weight", "global_img_feature_proj.bias"):
        del sd[k]
    _linear(g, sd, "global_img_feature_proj", E, H)
    return sd


def aoa_bu_state(seed: int, V: int, H: int, E: int, C: int = 2048):
    """AOAModelBU (aoamodel.py:1783-1806): Linear projector over C-d regions."""
    sd = aoa_decoder_state(seed, V, H, E, C=C)
    sd["img_projector.weight"] = sd["img_projector.weight"].reshape(H, C).clone()
    return sd


def bu_features(seed: int, n: int, n_region: int = 36, C: int = 2048):
    """Synthetic bottom-up region features, rand(n, 36, 2048)."""
    return torch.rand(n, n_region, C, generator=_gen(seed))


def resnet_state(seed: int, layers: Sequence[int] = (1, 1, 1, 1), randomize_bn: bool = True):
    """Bottleneck ResNet up to layer4 (resnet.py:143-239), incl. the unused fc head."""
    g = _gen(seed)
    sd = {}

    def conv(name, o, i, k):
        sd[name + ".weight"] = torch.randn(o, i, k, k, generator=g) * math.sqrt(2.0 / (o * k * k))

    def bn(name, c):
        if randomize_bn:
            sd[name + ".weight"] = torch.rand(c, generator=g) + 0.5
            sd[name + ".bias"] = torch.randn(c, generator=g) * 0.2
            sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.2
            sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
        else:
            sd[name + ".weight"] = torch.ones(c); sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c); sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("conv1", 64, 3, 7); bn("bn1", 64)
    inpl = 64
    for L, (planes, nb) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for b in range(nb):
            p = f"layer{L}.{b}"
            stride = 2 if (b == 0 and L > 1) else 1
            conv(p + ".conv1", planes, inpl, 1); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1); bn(p + ".bn3", planes * 4)
            if b == 0 and (stride != 1 or inpl != planes * 4):
                conv(p + ".downsample.0", planes * 4, inpl, 1); bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    sd["fc.weight"] = torch.zeros(1000, 2048); sd["fc.bias"] = torch.zeros(1000)
    return sd


def word_map(V: int) -> Dict[str, int]:
    """dataset/preparedataset.py:150-154 layout: <pad>=0, <unk>,<start>,<end> = V-3..V-1."""
    words = ["<pad>"] + [f"w{i}" for i in range(1, V - 3)] + ["<unk>", "<start>", "<end>"]
    return {w: i for i, w in enumerate(words)}


def stop_mask(V: int, every: int = 7) -> torch.Tensor:
    """Synthetic stop-word set: ids divisible by ``every`` + the specials."""
    m = torch.zeros(V, dtype=torch.bool)
    m[::every] = True
    m[[0, V - 3, V - 2, V - 1]] = True
    return m


def images(seed: int, n: int, size: int = 224) -> torch.Tensor:
    return torch.randn(n, 3, size, size, generator=_gen(seed))


def tokens(seed: int, T: int, V: int) -> List[int]:
    """[<start>] + T random content words."""
    t = torch.randint(1, V - 4, (T,), generator=_gen(seed)).tolist()
    return [V - 2] + t


def block_image_inputs(seed: int = 71):
    """Seeded heat-map-like inputs of the block_image fixture (oracle/make_golden.py golden_block_image)."""
    g = _gen(seed)
    heat = torch.randn(6, 3, 224, 224, generator=g) * torch.rand(6, 1, 224, 224, generator=g)
    heat[4, :, 8:16, 8:16] = 0                             # a patch of exact zeros
    heat[5] = heat[5].abs()                                # non-negative map
    small = torch.randn(3, 3, 32, 48, generator=g)
    return heat, small


def bbox_inputs(seed: int, n: int, n_boxes: int = 3, size: int = 224) -> torch.Tensor:
    """(n, n_boxes, 4) int boxes (x0, y0, x1, y1) inside a size x size image; the last box of image 0 is empty."""
    g = _gen(seed)
    lo = torch.randint(0, size - 40, (n, n_boxes, 2), generator=g)
    ext = torch.randint(8, 120, (n, n_boxes, 2), generator=g)
    b = torch.cat([lo, (lo + ext).clamp(max=size)], -1)
    b[0, -1, 2] = b[0, -1, 0]
    return b
