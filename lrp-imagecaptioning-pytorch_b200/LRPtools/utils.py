"""Mirror of the hot-path part of LRPtools/utils.py (reference :7-31, :55-64).

The heat-map post-processing the evaluation code feeds on (project / heatmap / graymap / gamma, utils.py:34-145) is
kept as host-side numpy like the reference's, without matplotlib: the two colour maps it uses ('seismic', 'gray') are
rebuilt as the 256-entry look-up tables matplotlib derives from their definitions.  ``visuallize_attention``
(utils.py:148-184: skimage up-sampling + an image overlay) stays out of scope.
"""
import numpy as np
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


# ---------------------------------------------------------------------------------- heat-map post-processing (host)
def project(X, output_range=(0, 1), absmax=None, input_is_postive_only=False):
    """utils.py:34-52.  X / max|X| (one maximum per leading index, zero maxima left alone; the division happens IN
    the caller's array, as in the reference), then [-1, 1] -> [0, 1] unless the input is positive only, clipped and
    stretched to ``output_range``."""
    amax = np.asarray(np.abs(X).max(axis=tuple(range(1, X.ndim))) if absmax is None else absmax)
    nz = amax != 0
    if nz.any():
        X[nz] /= amax[nz]
    unit = (X if input_is_postive_only else 0.5 * (X + 1)).clip(0, 1)
    lo, hi = output_range
    return lo + unit * (hi - lo)


def _lut(colors):
    """The 256-entry RGB table matplotlib builds for LinearSegmentedColormap.from_list(colors): evenly spaced colour
    nodes, linear interpolation at i / 255."""
    colors = np.asarray(colors, dtype=np.float64)
    at, nodes = np.linspace(0.0, 1.0, 256), np.linspace(0.0, 1.0, len(colors))
    return np.stack([np.interp(at, nodes, colors[:, c]) for c in range(3)], axis=1)


# 'seismic' = dark blue - blue - white - red - dark red (matplotlib _cm.py: _seismic_data); 'gray' = black - white
_CMAPS = {"seismic": _lut([(0, 0, 0.3), (0, 0, 1), (1, 1, 1), (1, 0, 0), (0.5, 0, 0)]), "gray": _lut([(0, 0, 0), (1, 1, 1)])}


def heatmap(X, cmap_type="seismic", reduce_op="sum", reduce_axis=-1, **kwargs):
    """utils.py:67-94: collapse ``reduce_axis`` (sum, or the signed entry of largest magnitude, ties to the positive
    one), ``project`` to 0..255, look the colour up; the collapsed axis becomes RGB."""
    table = _CMAPS.get(cmap_type)
    if table is None:
        raise NotImplementedError(f"colour map {cmap_type!r}: 'seismic' and 'gray' (the ones the reference uses) are built in")
    if reduce_op == "sum":
        flat = X.sum(axis=reduce_axis)
    elif reduce_op == "absmax":
        # Reference quirk kept (utils.py:76-81): it compares the positive extreme with the SIGNED negative extreme
        # (`abs_neg_max = -neg_max` is the minimum itself), which always holds, so "absmax" returns the maximum.
        top, bottom = X.max(axis=reduce_axis), X.min(axis=reduce_axis)
        flat = np.where(top >= bottom, top, -bottom)
    else:
        raise NotImplementedError()
    idx = project(flat, output_range=(0, 255), **kwargs).astype(np.int64).clip(0, 255)
    out_shape = list(X.shape)
    out_shape[reduce_axis] = 3
    return table[idx.ravel()].reshape(out_shape).astype(np.float32)


def graymap(X, **kwargs):
    """utils.py:97-98"""
    return heatmap(X, cmap_type="gray", **kwargs)


def gamma(X, gamma=0.7, minamp=0, maxamp=None):
    """utils.py:101-145: sign-preserving gamma curve around ``minamp``: y = sign(u) |u|^gamma with u = (X - minamp) /
    maxamp, scaled back by maxamp (default max|X|; an all-zero input is returned as it is)."""
    maxamp = np.abs(X).max() if maxamp is None else maxamp
    if maxamp == 0:
        return X
    u = (X - minamp) / maxamp
    return np.sign(u) * np.abs(u) ** gamma * maxamp + minamp
