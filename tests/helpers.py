"""Shared by the GPU parity tests, smoke() and bench.py: oracle state -> kernel argument dicts."""
import torch

import lrp_oracle as O
from lrpx import decoder as D

_GRID_RENAME = {"feat": "F", "st": "s"}
_AOA_RENAME = {"feat": "F"}


def gridtd_kernel_state(oracle_states, device):
    """oracle.gridtd_explainer_forward dicts (one per image) -> stacked kernel state."""
    keys = D.GRIDTD_IMAGE_KEYS + D.GRIDTD_STEP_KEYS + D.GRIDTD_STEP1_KEYS
    conv = [{k: st[_GRID_RENAME.get(k, k)] for k in keys} for st in oracle_states]
    return D.stack_states(conv, D.GRIDTD_IMAGE_KEYS, D.GRIDTD_STEP_KEYS, D.GRIDTD_STEP1_KEYS, device)


def aoa_kernel_state(oracle_states, device):
    keys = D.AOA_IMAGE_KEYS + D.AOA_STEP_KEYS + D.AOA_STEP1_KEYS
    conv = [{k: st[_AOA_RENAME.get(k, k)] for k in keys} for st in oracle_states]
    return D.stack_states(conv, D.AOA_IMAGE_KEYS, D.AOA_STEP_KEYS, D.AOA_STEP1_KEYS, device)


def adaptive_kernel_state(oracle_states, device):
    """oracle.adaptive_explainer_forward dicts (one per image) -> stacked kernel state (lrpx_adaptive_args)."""
    keys = D.ADAPTIVE_IMAGE_KEYS + D.ADAPTIVE_STEP_KEYS + D.ADAPTIVE_STEP1_KEYS
    conv = [{k: st[_GRID_RENAME.get(k, k)] for k in keys} for st in oracle_states]
    return D.stack_states(conv, D.ADAPTIVE_IMAGE_KEYS, D.ADAPTIVE_STEP_KEYS, D.ADAPTIVE_STEP1_KEYS, device)


def to_dev(d, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}


def gridtd_explainer_forward_ops(model, feat, tokens, quirk_double_bias_ih=True):
    """Step-by-step tensor-op restatement of ExplainGridTDAttention.get_hidden_parameters (reference
    gridTDmodel.py:941-1012) on a GridTDModel mirror: the checker of the fused explainer-forward kernels
    (lrpx_lstm_cell_f32 / lrpx_adaptive_attention_f32).  Test infrastructure; runs on CPU or GPU."""
    from models.gridTDmodel import _lstm_forward
    m = model
    B, P, C = feat.shape
    H, E = m.hidden_dim, m.embed_dim
    T = tokens.shape[1] - 1
    with torch.no_grad():
        avg = feat.mean(1)
        Wp = m.img_projector.weight.reshape(H, C)
        A_pre = feat @ Wp.t() + m.img_projector.bias
        A = A_pre.clamp(min=0)
        glob_pre = m.global_img_feature_proj(avg)
        glob = glob_pre.clamp(min=0)
        att = m.AdaAttention
        img_proj = att.W_v_proj(A)
        zeros = feat.new_zeros(B, H)
        h1, c1, h2, c2 = [zeros], [zeros], [zeros], [zeros]
        keys = ["x1", "x2", "g1", "i1", "f1", "g2", "i2", "f2", "st", "ctx", "ctx_hat", "alpha", "beta", "pred"]
        seq = {k: [] for k in keys}
        cell, L = m.AdaLSTM.lstm_cell, m.LanguageLSTM
        lb2 = L.bias_ih if quirk_double_bias_ih else L.bias_hh
        for t in range(T):
            emb = m.embedding(tokens[:, t])
            x1 = torch.cat((h2[t], glob, emb), dim=-1)
            h1n, c1n, g1, i1, f1 = _lstm_forward(x1, h1[t], c1[t], cell.weight_ih, cell.weight_hh, cell.bias_ih,
                                                 cell.bias_hh)
            s = torch.sigmoid(m.AdaLSTM.x_gate(x1) + m.AdaLSTM.h_gate(h1[t])) * torch.tanh(c1n)   # OLD h1 (:982)
            ctx_hat, ctx, alpha, beta = att.attend(A, img_proj, h1n, s)
            x2 = torch.cat((ctx_hat, h1n), dim=-1)
            h2n, c2n, g2, i2, f2 = _lstm_forward(x2, h2[t], c2[t], L.weight_ih, L.weight_hh, L.bias_ih, lb2)
            pred = m.fc(ctx_hat + h2n)
            for k, v in zip(keys, [x1, x2, g1, i1, f1, g2, i2, f2, s, ctx, ctx_hat, alpha, beta.squeeze(-1), pred]):
                seq[k].append(v)
            h1.append(h1n); c1.append(c1n); h2.append(h2n); c2.append(c2n)
        st = {k: torch.stack(v, 1).contiguous() for k, v in seq.items()}
        for k, v in (("h1", h1), ("c1", c1), ("h2", h2), ("c2", c2)):
            st[k] = torch.stack(v, 1).contiguous()
        st.update(feat=feat.contiguous(), avg=avg, A_pre=A_pre.contiguous(), A=A.contiguous(), glob_pre=glob_pre)
    return st


def aoa_explainer_forward_ops(model, feat, tokens, quirk_double_bias_ih=True):
    """Step-by-step tensor-op restatement of ExplainAOAAttention.get_hidden_parameters (reference aoamodel.py:999-1062)
    on an AOAModel mirror: the checker of the product's explainer forward (LSTM step kernel + batched attention); runs
    on any device."""
    import math
    from models.gridTDmodel import _lstm_forward
    m = model
    B, P, C = feat.shape
    H = m.hidden_dim
    T = tokens.shape[1] - 1
    nh, dk = m.num_head, H // m.num_head
    with torch.no_grad():
        Wp = m.img_projector.weight.reshape(H, C)
        A_pre = feat @ Wp.t() + m.img_projector.bias
        A = A_pre.clamp(min=0)
        glob = A.mean(1)
        key, value = m.decoder_k_proj(A), m.decoder_v_proj(A)
        kh = key.view(B, P, nh, dk).transpose(1, 2)
        vh = value.view(B, P, nh, dk).transpose(1, 2)
        mha = m.decoder_multihead_attention
        L = m.LanguageLSTM
        lb2 = L.bias_ih if quirk_double_bias_ih else L.bias_hh
        zeros = feat.new_zeros(B, H)
        h, c = [zeros], [zeros]
        keys = ["x", "g", "i", "f", "ctx", "caoa", "caoa_lin", "caoa_gate", "alpha", "pred"]
        seq = {k: [] for k in keys}
        for t in range(T):
            x = torch.cat((m.embedding(tokens[:, t]), glob), dim=-1)
            hn, cn, g, i, f = _lstm_forward(x, h[t], c[t], L.weight_ih, L.weight_hh, L.bias_ih, lb2)
            q = mha.q_proj(hn).view(B, nh, 1, dk)
            alpha = torch.softmax(torch.matmul(q, kh.transpose(-2, -1)) / math.sqrt(dk), dim=-1)   # (B,nh,1,P)
            ctx = torch.matmul(alpha, vh).transpose(1, 2).reshape(B, H)
            gate = m.decoder_aoa_linear_gate(hn)
            lin = m.decoder_aoa_linear(ctx)
            caoa = torch.sigmoid(gate) * lin
            pred = m.fc(caoa + hn)
            for k, v in zip(keys, [x, g, i, f, ctx, caoa, lin, gate, alpha.squeeze(2), pred]):
                seq[k].append(v)
            h.append(hn); c.append(cn)
        st = {k: torch.stack(v, 1).contiguous() for k, v in seq.items()}
        st["h"], st["c"] = torch.stack(h, 1).contiguous(), torch.stack(c, 1).contiguous()
        st.update(feat=feat.contiguous(), A_pre=A_pre.contiguous(), A=A.contiguous(), glob=glob, key=key,
                  value=value.contiguous())
    return st


def gridtd_grad_kernel_state(oracle_states, device):
    """oracle.gridtd_explainer_forward(..., gradient=True) dicts -> stacked state of lrpx_gridtd_grad_args"""
    step = D.GRIDTD_STEP_KEYS + ["o1", "o2", "sg"]
    conv = [{k: st[_GRID_RENAME.get(k, k)] for k in D.GRIDTD_IMAGE_KEYS + step + D.GRIDTD_STEP1_KEYS} for st in oracle_states]
    return D.stack_states(conv, D.GRIDTD_IMAGE_KEYS, step, D.GRIDTD_STEP1_KEYS, device)


def aoa_grad_kernel_state(oracle_states, device):
    step = D.AOA_STEP_KEYS + ["f", "o", "caoa_gate"]
    conv = [{k: st[_AOA_RENAME.get(k, k)] for k in D.AOA_IMAGE_KEYS + step + D.AOA_STEP1_KEYS} for st in oracle_states]
    return D.stack_states(conv, D.AOA_IMAGE_KEYS, step, D.AOA_STEP1_KEYS, device)
