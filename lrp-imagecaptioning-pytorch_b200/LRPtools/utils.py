"""Mirror of the hot-path part of LRPtools/utils.py (reference :7-31, :55-64).

The visualisation helpers of the reference (project/heatmap/graymap/gamma/visuallize_attention,
utils.py:34-184) depend on matplotlib/skimage and are out of scope.
"""
import torch

from lrpx import ops as _ops

LOWEST = -1
HIGHEST = 1
EPSILON = 0.01
Z_EPSILON = 1e-7
LOGIT_BETA = 4
RELEVANCE_RECT = -1e-6
ALPHA = 1.
BETA = 0.


def safe_divide(numerator, divisor):
    """utils.py:16-18.  Pure tensor expression (device of the inputs); the fused kernels apply the same guard."""
    return numerator / (divisor + Z_EPSILON * (divisor == 0).float())


def lrp_backward(_input, layer, relevance_output):
    """utils.py:21-31 for the sign-split conv nets of lrp_modules (PosNetConv / NegNetConv):
    s = R / guard(layer(_input));  returns _input * d layer/d input (s).  Runs K1 + K2 of liblrpx."""
    from . import lrp_modules
    if not isinstance(layer, (lrp_modules.PosNetConv, lrp_modules.NegNetConv)):
        raise NotImplementedError("lrp_backward is implemented for PosNetConv / NegNetConv layers")
    return layer.lrp_backward(_input, relevance_output)


def normalize_relevance(X, dim=-1, temperature=1):
    """utils.py:55-64: X / max|X| (row-wise) * temperature + (temperature if temperature > 1 else 1)."""
    if dim not in (-1, X.dim() - 1):
        X = X.transpose(dim, -1)
        return _ops.normalize_relevance(X, temperature).transpose(dim, -1)
    return _ops.normalize_relevance(X, temperature)
