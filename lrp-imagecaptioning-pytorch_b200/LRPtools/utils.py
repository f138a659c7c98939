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
    """utils.py:34-52: scale by max|X| (per leading index), map [-1,1] -> [0,1] unless the input is positive only,
    clip, stretch to ``output_range``.  Like the reference it divides ``X`` in place."""
    if absmax is None:
        absmax = np.max(np.abs(X), axis=tuple(range(1, len(X.shape))))
    absmax = np.asarray(absmax)
    mask = absmax != 0
    if mask.sum() > 0:
        X[mask] /= absmax[mask]
    if input_is_postive_only is False:
        X = (X + 1) / 2
    X = X.clip(0, 1)
    return output_range[0] + (X * (output_range[1] - output_range[0]))


def _lut(nodes):
    """256-entry RGB table of a matplotlib LinearSegmentedColormap.from_list(colors): linear interpolation between
    evenly spaced nodes at i / 255."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.linspace(0.0, 1.0, 256)
    xp = np.linspace(0.0, 1.0, len(nodes))
    return np.stack([np.interp(x, xp, nodes[:, c]) for c in range(3)], axis=1)


_CMAPS = {"seismic": _lut([(0, 0, 0.3), (0, 0, 1), (1, 1, 1), (1, 0, 0), (0.5, 0, 0)]),      # matplotlib _cm.py: _seismic_data
          "gray": _lut([(0, 0, 0), (1, 1, 1)])}


def heatmap(X, cmap_type="seismic", reduce_op="sum", reduce_axis=-1, **kwargs):
    """utils.py:67-94: reduce the channel axis (sum, or the entry of largest magnitude), project to 0..255, colour."""
    if cmap_type not in _CMAPS:
        raise NotImplementedError(f"colour map {cmap_type!r}: only 'seismic' and 'gray' (what the reference uses) are built in")
    tmp = X
    shape = tmp.shape
    if reduce_op == "sum":
        tmp = tmp.sum(axis=reduce_axis)
    elif reduce_op == "absmax":
        pos_max = tmp.max(axis=reduce_axis)
        neg_max = (-tmp).max(axis=reduce_axis)
        abs_neg_max = -neg_max
        tmp = np.select([pos_max >= abs_neg_max, pos_max < abs_neg_max], [pos_max, neg_max])
    else:
        raise NotImplementedError()
    tmp = project(tmp, output_range=(0, 255), **kwargs).astype(np.int64)
    tmp = _CMAPS[cmap_type][tmp.flatten().clip(0, 255)]
    shape = list(shape)
    shape[reduce_axis] = 3
    return tmp.reshape(shape).astype(np.float32)


def graymap(X, **kwargs):
    """utils.py:97-98"""
    return heatmap(X, cmap_type="gray", **kwargs)


def gamma(X, gamma=0.7, minamp=0, maxamp=None):
    """utils.py:101-145: gamma correction of positive and negative values separately around ``minamp``."""
    if maxamp is None:
        maxamp = np.abs(X).max()
    if maxamp == 0:
        return X
    Y = np.zeros_like(X)
    X = X - minamp
    X = X / maxamp
    i_pos = X >= 0
    if i_pos.sum() > 0:
        Y[i_pos] = X[i_pos] ** gamma
    i_neg = np.invert(i_pos)
    if i_neg.sum() > 0:
        Y[i_neg] = -(-X[i_neg]) ** gamma
    Y *= maxamp
    Y += minamp
    return Y
