"""Mirror of LRPtools/lrp_modules.py: the per-layer relevance rule objects.

Same class names, same ``propagate_relevance(module, relevance_input, relevance_output, lrp_method,
lrp_params=None)`` signature, same tuple arity/order as torch's ``grad_input`` (reference :29-37,
:157-170), same assertions (NaN/Inf ``AssertionError``, ``NotImplementedError`` for a non-alpha-beta
conv, ``ValueError`` for an unknown layer).  The arithmetic is liblrpx.so (fp32 parity path).
"""
import torch
import torch.nn as nn

from lrpx import ops
import LRPtools.utils as util
from models.resnet import Add as resAdd
from models.resnet import Flatten as resFlatten


def _check_finite(R):
    s = R.sum()
    assert not torch.isnan(s)
    assert not torch.isinf(s)


def _like_grad_input(R, relevance_input, position=0):
    """Returns a tuple with the arity of ``relevance_input`` (torch's grad_input) holding R at ``position``."""
    if relevance_input is None:
        return (R,)
    out = list(relevance_input)
    out[position] = R
    return tuple(out)


class Linear:
    """reference :9-37 — epsilon rule in GEMM form (zeros of the input count as -1e-6, Q9)."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        ignore_bias = (lrp_params or {}).get("ignore_bias", True)
        input_ = module.input[0]
        R = ops.linear_epsilon(input_.reshape(-1, input_.shape[-1]), module.weight.detach(),
                               None if module.bias is None else module.bias.detach(),
                               relevance_output[0].reshape(-1, module.weight.shape[0]), ignore_bias).view_as(input_)
        assert R.shape == input_.shape
        _check_finite(R)
        if relevance_input is not None and len(relevance_input) == 3:
            return relevance_input[0], R, relevance_input[2]
        return _like_grad_input(R, relevance_input, 0)


class ReLU:
    """reference :39-54."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        if lrp_method == 'identity':
            return (relevance_output[0],)
        R = ops.relu_mask(module.input[0], relevance_output[0])
        _check_finite(R)
        return (R,)


class _SignedNetConv(nn.Module):
    """PosNetConv / NegNetConv (reference :56-114) without cloning or re-initialising conv modules: the
    W+/W- and a+/a- split happens inside the kernel."""
    NET = ops.NET_POS

    def __init__(self, conv, ignorebias):
        super().__init__()
        self.conv = conv
        self.ignorebias = ignorebias

    def _bias(self):
        return None if (self.ignorebias or self.conv.bias is None) else self.conv.bias.detach()

    def forward(self, x):
        c = self.conv
        zero = torch.zeros(1, device=x.device)
        shp = ops.conv_shape(x, c.weight, c.stride, c.padding, c.dilation)
        P, Q = ops.conv_out_hw(shp)
        r = zero.expand(shp.n, shp.cout, P, Q).contiguous()
        _, z = ops.conv_rule_s(x, c.weight.detach(), self._bias(), r, c.stride, c.padding, c.dilation, self.NET, True)
        return z

    def lrp_backward(self, x, relevance_output):
        c = self.conv
        s = ops.conv_rule_s(x, c.weight.detach(), self._bias(), relevance_output, c.stride, c.padding, c.dilation,
                            self.NET)
        return ops.conv_rule_rin(x, c.weight.detach(), s, c.stride, c.padding, c.dilation, self.NET)


class PosNetConv(_SignedNetConv):
    NET = ops.NET_POS


class NegNetConv(_SignedNetConv):
    NET = ops.NET_NEG


class Conv2d:
    """reference :116-170 — alpha-beta rule: alpha * R(pos-net) - beta * R(neg-net)."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        lrp_params = lrp_params or {}
        ignore_bias = lrp_params.get("ignore_bias", True)
        input_ = module.input[0]
        if lrp_method == "alpha_beta":
            if module.groups != 1:
                raise NotImplementedError("grouped convolutions are not supported")
            R = ops.conv_alpha_beta(input_.detach(), module.weight.detach(),
                                    None if module.bias is None else module.bias.detach(), relevance_output[0].detach(),
                                    module.stride, module.padding, module.dilation, lrp_params["alpha"],
                                    lrp_params["beta"], ignore_bias)
        else:
            raise NotImplementedError('Only adopt alpha 1 rule for conv layer')
        assert R.shape == input_.shape
        _check_finite(R)
        return _like_grad_input(R, relevance_input, 0)


class Pool2d:
    """reference :172-195 — max: winner-take-all, avg: proportional."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        input_ = module.input[0]
        if type(module) == nn.MaxPool2d:
            if module.dilation not in (1, (1, 1)) or module.ceil_mode:
                raise NotImplementedError("max-pool with dilation / ceil_mode is not supported")
            R = ops.maxpool_wta(input_.detach(), relevance_output[0].detach(), module.kernel_size, module.stride,
                                module.padding)
        elif type(module) == nn.AvgPool2d:
            if module.ceil_mode or not module.count_include_pad:
                raise NotImplementedError("avg-pool with ceil_mode / count_include_pad=False is not supported")
            R = ops.avgpool_prop(input_.detach(), relevance_output[0].detach(), module.kernel_size, module.stride,
                                 module.padding)
        else:
            raise ValueError(type(module))
        _check_finite(R)
        return (R,)


class BatchNorm2d:
    """reference :197-219 — R = |x w| / (|x w| + |b|) * R_out (Q12)."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        if lrp_method == 'identity':
            R = relevance_output[0]
        else:
            R = ops.bn_absratio(module.input[0].detach(), relevance_output[0].detach(), module.running_mean,
                                module.running_var, module.weight.detach(), module.bias.detach(), module.eps)
        _check_finite(R)
        assert R.sum() != 0
        if relevance_input is not None and len(relevance_input) == 3:
            return R, relevance_input[1], relevance_input[2]
        return (R,)


class BatchNorm1d(BatchNorm2d):
    """reference :221-246 — same expression (the reference reuses the 2-d broadcasting)."""


class Dropout:
    """reference :248-254."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        if relevance_input is None:
            return (relevance_output[0],)
        assert ((relevance_output[0] - relevance_input[0]).abs().max() < 1e-7).cpu().item() == 1
        return relevance_input


class Add:
    """reference :256-280 — epsilon-stabilised proportional split (Q14)."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        R1, R2 = ops.add_split(module.input[0].detach(), module.input[1].detach(), relevance_output[0].detach())
        _check_finite(R1)
        _check_finite(R2)
        return R1, R2


class Flatten:
    """reference :282-291."""

    def propagate_relevance(self, module, relevance_input, relevance_output, lrp_method, lrp_params=None):
        R = relevance_output[0].detach().reshape(module.input[0].size())
        _check_finite(R)
        return (R,)


def compute_lrp_sum(sum_output, sum_input, relevance_sum_output, dim=-1):
    """reference :293-304 (unused helper; tensor expression)."""
    assert (sum_output == torch.sum(sum_input, dim=dim)).all()
    fea_dim = sum_input.size()[-1]
    relevance = relevance_sum_output.unsqueeze(-1).repeat(1, 1, 1, fea_dim)
    out = sum_output.unsqueeze(-1).repeat(1, 1, 1, fea_dim)
    out = out.masked_fill(out == 0, 1 / fea_dim)
    return relevance * sum_input / (out + util.EPSILON * out.sign())


def compute_lrp_mean(mean_output, mean_input, relevance_mean_output, dim=-1):
    """reference :306-319 (unused helper; tensor expression)."""
    assert (mean_output == torch.mean(mean_input, dim=dim)).all()
    fea_dim = mean_input.size()[-1]
    rep = [1] * len(mean_input.shape)
    rep[-1] *= fea_dim
    relevance = relevance_mean_output.unsqueeze(-1).repeat(rep)
    out = mean_input.sum(dim=dim).unsqueeze(-1).repeat(rep)
    out = out.masked_fill(out == 0, 1 / fea_dim)
    return relevance * mean_input / (out + util.EPSILON * out.sign())


def get_lrp_module(module):
    """reference :321-341."""
    try:
        lrp_module_class = {
            nn.Linear: Linear,
            nn.ReLU: ReLU,
            nn.Conv2d: Conv2d,
            nn.MaxPool2d: Pool2d,
            nn.AvgPool2d: Pool2d,
            nn.BatchNorm2d: BatchNorm2d,
            nn.BatchNorm1d: BatchNorm1d,
            nn.Dropout: Dropout,
            nn.Dropout2d: Dropout,
            resFlatten: Flatten,
            resAdd: Add,
        }[type(module)]
    except KeyError:
        raise ValueError("Layer type {} not known.".format(type(module)))
    return lrp_module_class()
