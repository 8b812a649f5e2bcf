"""Drop-in replacement for the reference's ``modules_realnvp.py`` (B200-native compute).

Same class names, constructor signatures, attribute names and ``state_dict``
layout as the reference (modules_realnvp.py:36-370), so that
checkpoints and optimizer states interchange and ``flow_realnvp.RealNVP`` /
``train.py`` run unchanged.  The parameters are created with the same torch
constructors in the same order, which also makes the initialisation
bit-identical under a given seed.  What differs is who does the arithmetic: no
module here calls a torch operator on its hot path -- forward, inverse and
backward of a coupling are launches of the sm_100a kernels behind the C-ABI
(include/rnvp.h), driven through :mod:`rnvp_engine`.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn

import rnvp_engine as _eng


def _hps_check(hps):
    """Only the configuration the shipped caller selects (train.py:121-128) is built natively."""
    bad = [k for k in ("bottleneck", "skip", "weight_norm", "coupling_bn") if not getattr(hps, k)]
    if bad or hps.res_blocks < 1:
        raise NotImplementedError(
            "rnvp-b200 implements bottleneck=skip=weight_norm=coupling_bn=True with res_blocks>=1 "
            f"(the configuration train.py constructs); got {bad or 'res_blocks=0'}")


class WeightNormConv2d(nn.Module):
    """Weight-normalised conv: parameters ``conv.{bias,weight_g,weight_v}`` (modules_realnvp.py:36-71)."""

    def __init__(self, in_dim, out_dim, kernel_size, stride=1, padding=0,
                 bias=True, weight_norm=True, scale=False):
        super().__init__()
        conv = nn.Conv2d(in_dim, out_dim, kernel_size, stride=stride, padding=padding, bias=bias)
        if weight_norm:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                conv = nn.utils.weight_norm(conv)          # old-style: registers weight_g / weight_v
            if not scale:                                   # frozen magnitude, still a Parameter
                conv.weight_g.data = torch.ones_like(conv.weight_g.data)
                conv.weight_g.requires_grad = False
        self.conv = conv

    def forward(self, x):
        raise NotImplementedError(
            "WeightNormConv2d only holds parameters here; it is evaluated inside a coupling's fused "
            "s/t-network kernels (CheckerboardAffineCoupling / ChannelwiseAffineCoupling)")


class ResidualBlock(nn.Module):
    """BN-ReLU-1x1-BN-ReLU-3x3-BN-ReLU-1x1 bottleneck with identity skip (modules_realnvp.py:73-114)."""

    def __init__(self, dim, bottleneck, weight_norm):
        super().__init__()
        if not bottleneck:
            raise NotImplementedError("rnvp-b200 builds the bottleneck residual block only")
        self.in_block = nn.Sequential(nn.BatchNorm2d(dim), nn.ReLU())
        self.res_block = nn.Sequential(
            WeightNormConv2d(dim, dim, (1, 1), stride=1, padding=0, bias=False, weight_norm=weight_norm, scale=False),
            nn.BatchNorm2d(dim), nn.ReLU(),
            WeightNormConv2d(dim, dim, (3, 3), stride=1, padding=1, bias=False, weight_norm=weight_norm, scale=False),
            nn.BatchNorm2d(dim), nn.ReLU(),
            WeightNormConv2d(dim, dim, (1, 1), stride=1, padding=0, bias=True, weight_norm=weight_norm, scale=True))

    def forward(self, x):
        raise NotImplementedError("ResidualBlock is evaluated inside a coupling's fused s/t-network kernels")


class ResidualModule(nn.Module):
    """The s/t network: in conv, ``res_blocks`` bottleneck blocks with 1x1 skip taps, out conv
    (modules_realnvp.py:116-194)."""

    def __init__(self, in_dim, dim, out_dim, res_blocks, bottleneck, skip, weight_norm):
        super().__init__()
        if res_blocks < 1 or not skip or not bottleneck:
            raise NotImplementedError("rnvp-b200 builds res_blocks>=1, skip=True, bottleneck=True only")
        self.res_blocks, self.skip = res_blocks, skip
        # construction order = the reference's, so seeds reproduce its initialisation
        self.in_block = WeightNormConv2d(in_dim, dim, (3, 3), stride=1, padding=1, bias=True,
                                         weight_norm=weight_norm, scale=False)
        self.core_block = nn.ModuleList(ResidualBlock(dim, bottleneck, weight_norm) for _ in range(res_blocks))
        self.out_block = nn.Sequential(
            nn.BatchNorm2d(dim), nn.ReLU(),
            WeightNormConv2d(dim, out_dim, (1, 1), stride=1, padding=0, bias=True, weight_norm=weight_norm, scale=True))
        self.in_skip = WeightNormConv2d(dim, dim, (1, 1), stride=1, padding=0, bias=True,
                                        weight_norm=weight_norm, scale=True)
        self.core_skips = nn.ModuleList(
            WeightNormConv2d(dim, dim, (1, 1), stride=1, padding=0, bias=True, weight_norm=weight_norm, scale=True)
            for _ in range(res_blocks))

    def forward(self, x):
        raise NotImplementedError("ResidualModule is evaluated inside a coupling's fused s/t-network kernels")


class AbstractCoupling(nn.Module):
    """Shared bookkeeping of the two coupling kinds (modules_realnvp.py:196-237)."""

    _KIND = -1

    def __init__(self, mask_config, hps):
        super().__init__()
        _hps_check(hps)
        self.mask_config = mask_config
        self.res_blocks = hps.res_blocks
        self.bottleneck = hps.bottleneck
        self.skip = hps.skip
        self.weight_norm = hps.weight_norm
        self.coupling_bn = hps.coupling_bn
        self._engine = None
        self._math = None                 # None = package default; set through set_math()

    def build_mask(self, size, config=1.):
        """(1,1,size,size) float mask, mask[i,j] = (config+i+j) mod 2 (modules_realnvp.py:211-226)."""
        idx = np.add.outer(np.arange(size), np.arange(size))
        return torch.from_numpy(np.mod(config + idx, 2).astype("float32").reshape(1, 1, size, size))

    def batch_stat(self, x):
        """Per-channel mean and biased variance over (N,H,W), keepdim (modules_realnvp.py:228-237).
        Utility kept for API compatibility; the kernels compute these statistics themselves."""
        mean = x.mean(dim=(0, 2, 3), keepdim=True)
        return mean, ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)

    # nn.Module._apply runs for .to()/.cuda()/.float(): parameter storage may move
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine.dirty = True
        return out

    def _shape(self):
        raise NotImplementedError

    def _all_engines(self):
        out = [self._engine] if self._engine is not None else []
        return out + list(self.__dict__.get("_engines_by_size", {}).values())

    def set_math(self, mode):
        """'tf32' / 'fp32' arithmetic tier for this module's stand-alone calls (None = default)."""
        import rnvp_cabi
        self._math = None if mode is None else {"fp32": rnvp_cabi.MATH_FP32, "tf32": rnvp_cabi.MATH_TF32}[mode]
        for e in self._all_engines():
            e.set_math(_eng._DEFAULT_MATH if self._math is None else self._math)

    def _own_engine(self):
        if self._engine is None:
            c, s, d = self._shape()
            self._engine = _eng.Engine.for_coupling(self._KIND, c, s, d, int(self.mask_config), self.res_blocks, self,
                                                    math=self._math)
        return self._engine

    def forward(self, x, reverse=False):
        """Returns ``(transformed x, log_diag_J)`` like the reference (modules_realnvp.py:264, 324)."""
        self._check_input(x)
        eng = self._own_engine()
        if reverse:
            with torch.no_grad():
                y = eng.coupling_inverse(0, x, self.training)
            # the reference returns log_rescale here; its callers discard it (flow_realnvp.py:203)
            return y, None
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _eng.CouplingFn.apply(_eng.grad_anchor(x.device), x, eng, self.training)
        y, logj, _ = eng.coupling_forward(0, x, self.training)
        return y, logj

    def _check_input(self, x):
        c, s, _ = self._shape()
        if x.dim() != 4 or x.shape[1] != c or (s is not None and (x.shape[2] != s or x.shape[3] != s)):
            raise ValueError(f"expected input of shape (B,{c},{s},{s}), got {tuple(x.shape)}")


class CheckerboardAffineCoupling(AbstractCoupling):
    """Affine coupling with a checkerboard mask (modules_realnvp.py:239-302)."""

    _KIND = 0

    def __init__(self, in_out_dim, mid_dim, size, mask_config, hps):
        super().__init__(mask_config, hps)
        self.mask = self.build_mask(size, config=mask_config)     # kept for API parity; kernels derive it
        self._dims = (in_out_dim, size, mid_dim)
        self.scale = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.scale_shift = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.in_bn = nn.BatchNorm2d(in_out_dim)
        self.block = nn.Sequential(
            nn.ReLU(),
            ResidualModule(2 * in_out_dim + 1, mid_dim, 2 * in_out_dim,
                           self.res_blocks, self.bottleneck, self.skip, self.weight_norm))
        self.out_bn = nn.BatchNorm2d(in_out_dim, affine=False)

    def _shape(self):
        return self._dims


class ChannelwiseAffineCoupling(AbstractCoupling):
    """Affine coupling that transforms one half of the channels (modules_realnvp.py:304-370)."""

    _KIND = 1

    def __init__(self, in_out_dim, mid_dim, mask_config, hps):
        super().__init__(mask_config, hps)
        self._dims = (in_out_dim, None, mid_dim)
        self.scale = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.scale_shift = nn.Parameter(torch.zeros(1), requires_grad=True)
        self.in_bn = nn.BatchNorm2d(in_out_dim // 2)
        self.block = nn.Sequential(
            nn.ReLU(),
            ResidualModule(in_out_dim, mid_dim, in_out_dim,
                           self.res_blocks, self.bottleneck, self.skip, self.weight_norm))
        self.out_bn = nn.BatchNorm2d(in_out_dim // 2, affine=False)

    def _shape(self):
        return self._dims

    def _own_engine(self):
        raise RuntimeError("internal: channelwise engines are keyed by spatial size")   # pragma: no cover

    def forward(self, x, reverse=False):
        # the spatial size is not a constructor argument of this class: plans are keyed by it
        self._check_input(x)
        s = x.shape[2]
        engines = self.__dict__.setdefault("_engines_by_size", {})
        if s not in engines:
            c, _, d = self._dims
            engines[s] = _eng.Engine.for_coupling(1, c, s, d, int(self.mask_config), self.res_blocks, self, math=self._math)
        eng = engines[s]
        if reverse:
            with torch.no_grad():
                return eng.coupling_inverse(0, x, self.training), None
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _eng.CouplingFn.apply(_eng.grad_anchor(x.device), x, eng, self.training)
        y, logj, _ = eng.coupling_forward(0, x, self.training)
        return y, logj

    def _apply(self, fn, *a, **k):
        out = nn.Module._apply(self, fn, *a, **k)
        for e in self.__dict__.get("_engines_by_size", {}).values():
            e.dirty = True
        return out
