"""Packing helpers between the explainers' saved state (get_hidden_parameters, gridTDmodel.py:933-1012 /
aoamodel.py:990-1062 / adaptiveattention.py:626-677) and the batched decoder-relevance kernels (lrpx_gridtd_args / lrpx_aoa_args / lrpx_adaptive_args)."""
from typing import Dict, List, Sequence

import torch

# per-image tensors / per-step tensors (leading dim T) / per-step-plus-one tensors (leading dim T+1)
GRIDTD_IMAGE_KEYS = ["feat", "avg", "A_pre", "A", "glob_pre"]
GRIDTD_STEP_KEYS = ["x1", "x2", "g1", "i1", "f1", "g2", "i2", "f2", "st", "ctx", "ctx_hat", "alpha", "beta", "pred"]
GRIDTD_STEP1_KEYS = ["h1", "c1", "h2", "c2"]
AOA_IMAGE_KEYS = ["feat", "A_pre", "A", "glob", "value"]
AOA_STEP_KEYS = ["x", "g", "i", "ctx", "caoa", "caoa_lin", "alpha", "pred"]
AOA_STEP1_KEYS = ["h", "c"]
ADAPTIVE_IMAGE_KEYS = ["feat", "avg", "z_proj", "A", "z_glob"]
ADAPTIVE_STEP_KEYS = ["x", "g", "i", "f", "st", "ctx", "ctx_hat", "alpha", "beta", "pred"]
ADAPTIVE_STEP1_KEYS = ["h", "c"]


def _gate_rows(w_ih, w_hh):
    """Rows of the cell-candidate gate g (chunk 2 of i,f,g,o) of [W_ih | W_hh] (gridTDmodel.py:1019-1024)."""
    return torch.cat([w_ih.chunk(4, 0)[2], w_hh.chunk(4, 0)[2]], 1).contiguous()


def gridtd_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    H = sd["fc.weight"].shape[1]
    return {
        "W_g1": _gate_rows(sd["AdaLSTM.lstm_cell.weight_ih"], sd["AdaLSTM.lstm_cell.weight_hh"]),
        "W_g2": _gate_rows(sd["LanguageLSTM.weight_ih"], sd["LanguageLSTM.weight_hh"]),
        "W_fc": sd["fc.weight"].contiguous(),
        "W_glob": sd["global_img_feature_proj.weight"].contiguous(),
        "W_proj": sd["img_projector.weight"].reshape(H, -1).contiguous(),
    }


def aoa_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    H = sd["fc.weight"].shape[1]
    return {
        "W_g": _gate_rows(sd["LanguageLSTM.weight_ih"], sd["LanguageLSTM.weight_hh"]),
        "W_fc": sd["fc.weight"].contiguous(),
        "W_aoa": sd["decoder_aoa_linear.weight"].contiguous(),
        "W_v": sd["decoder_v_proj.weight"].contiguous(),
        "W_proj": sd["img_projector.weight"].reshape(H, -1).contiguous(),
    }


def adaptive_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """adaptiveattention.py:685-687 (gate-g rows of the AdaLSTM), :523, :746, :763."""
    H = sd["fc.weight"].shape[1]
    return {
        "W_g": _gate_rows(sd["AdaLSTM.lstm_cell.weight_ih"], sd["AdaLSTM.lstm_cell.weight_hh"]),
        "W_fc": sd["fc.weight"].contiguous(),
        "W_glob": sd["global_img_feature_proj.weight"].contiguous(),
        "W_proj": sd["img_projector.weight"].reshape(H, -1).contiguous(),
    }


def stack_states(states: Sequence[Dict[str, torch.Tensor]], image_keys: List[str], step_keys: List[str],
                 step1_keys: List[str], device) -> Dict[str, torch.Tensor]:
    """Stacks per-image state dicts over a leading batch dim, zero-padding the step dim to the longest caption."""
    T = max(int(s[step_keys[0]].shape[0]) for s in states)
    out = {}
    for k in image_keys:
        out[k] = torch.stack([s[k].to(device=device, dtype=torch.float32) for s in states]).contiguous()
    for keys, extra in ((step_keys, 0), (step1_keys, 1)):
        for k in keys:
            rows = []
            for s in states:
                v = s[k].to(device=device, dtype=torch.float32)
                pad = T + extra - v.shape[0]
                if pad:
                    v = torch.cat([v, v.new_zeros((pad,) + tuple(v.shape[1:]))])
                rows.append(v)
            out[k] = torch.stack(rows).contiguous()
    return out
