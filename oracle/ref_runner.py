"""Drives the UNMODIFIED reference (through ref_shim) on the host CPU for bench.py's reference arm and for the parity
figures it prints: the reference's own ExplainGridTDAttention (gridTDmodel.py:705-1211) and its VGG16 encoder with
add_lrp / compute_lrp (lrp_wrapper.py:37-87), on the same seeded weights, images and captions as the CUDA arm.
TEST / BENCH INFRASTRUCTURE ONLY — never imported by the product path."""
import argparse
import contextlib
import io
import os
import warnings

import torch

import ref_shim


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def available():
    return ref_shim.reference_available()


class GridTDReference:
    """Reference GridTDModel (VGG16 encoder) + ExplainGridTDAttention with the given state dicts."""

    def __init__(self, dec_state, vgg_state, V, H, E, word_map):
        with _quiet():
            self.ns = ns = ref_shim.load_reference()
            model = ns.gridTDmodel.GridTDModel(E, H, V, "vgg16")
        model.load_state_dict(dec_state, strict=False)
        model.img_encoder.encoder.load_state_dict(vgg_state)
        model.eval()
        os.makedirs("/tmp/lrpx_ref", exist_ok=True)
        args = argparse.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                                  save_path="/tmp/lrpx_ref", dataset="syn", weight="")
        self.ex = ns.gridTDmodel.ExplainGridTDAttention(args, word_map, model=model)
        self.model = model
        with _quiet():
            ns.lrp_wrapper.add_lrp(model.img_encoder.encoder)          # once (the reference stacks hooks per call, Q2)

    def set_image(self, img, tokens):
        """img (1,3,224,224), tokens = [<start>, w1, ..., wT] -> runs get_hidden_parameters (gridTDmodel.py:933-1012)"""
        ex = self.ex
        ex.preprocess_img = lambda p: img
        words = list(tokens[1:])
        self.model.beam_search = lambda *a, **k: ([" ".join(f"w{t}" for t in words)], words)
        with torch.no_grad(), _quiet():
            ex.get_hidden_parameters("x")

    def explain_word(self, img, t):
        """-> (heat (1,3,224,224), r_words (t+1,), r_img_feature (1,512,14,14)): explain_caption_wordt + explain_cnn
        with a fresh sample per call (Q1 accumulation disabled, as on the CUDA arm)."""
        ex = self.ex
        with _quiet():
            with torch.no_grad():
                r_feat, r_words = ex.explain_caption_wordt(t)
            ex.img = img.detach().clone()
            heat = ex.explain_cnn(r_feat)
        return heat.detach(), r_words.detach(), r_feat.detach()


class VggReference:
    """Reference VGG16 features[0:-1] with add_lrp; compute_lrp(sample, target) (lrp_wrapper.py:63-87)."""

    def __init__(self, vgg_state):
        with _quiet():
            self.ns = ns = ref_shim.load_reference()
            net = ns.vgg.vgg16(pretrained=False).features[0:-1]
        net.load_state_dict(vgg_state)
        net.eval()
        with _quiet():
            ns.lrp_wrapper.add_lrp(net)
        self.net = net

    def compute_lrp(self, x, target):
        with _quiet():
            return self.net.compute_lrp(x.detach().clone(), target=target).detach()
