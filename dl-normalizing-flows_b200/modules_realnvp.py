"""Drop-in replacement for the reference's ``modules_realnvp.py`` (B200-native compute).

Same class names, constructor signatures, attribute names and ``state_dict``
layout as the reference (modules_realnvp.py:36-370), so that
checkpoints and optimizer states interchange and ``flow_realnvp.RealNVP`` /
``train.py`` run unchanged.  The parameters are created with the same torch
constructors in the same order, which also makes the initialisation
bit-identical under a given seed.  What differs is who does the arithmetic:

* the configuration the shipped caller constructs (train.py:121-128:
  ``bottleneck = skip = weight_norm = coupling_bn = True``, ``res_blocks >= 1``)
  runs forward, inverse and backward of a coupling as launches of the sm_100a
  kernels behind the C-ABI (include/rnvp.h), driven through :mod:`rnvp_engine`
  -- no torch operator on that path and no CPU fallback;
* the hyper-parameter branches the caller never selects and
  the stand-alone ``forward`` of the helper modules (``WeightNormConv2d``,
  ``ResidualBlock``, ``ResidualModule``: inside a coupling they are evaluated by
  the fused kernels, never called) are API-completeness paths: they compose
  ATen operators on the CUDA device, with autograd.  They still refuse CPU
  tensors -- this package has no CPU path.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn

import rnvp_engine as _eng


def _native_hps(hps) -> bool:
    """True for the configuration the sm_100a kernels implement (the one train.py:121-128 constructs)."""
    return bool(hps.bottleneck and hps.skip and hps.weight_norm and hps.coupling_bn and hps.res_blocks >= 1)


def _cuda_only(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor -- this package has no CPU path")
    return x


class WeightNormConv2d(nn.Module):
    """Conv2d with (old-style) weight normalisation: parameters ``conv.{bias,weight_g,weight_v}``, or a plain
    ``conv.{weight,bias}`` when ``weight_norm=False`` (modules_realnvp.py:36-71)."""

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
        """Stand-alone call (ATen on the CUDA device).  Inside a coupling this module only holds parameters."""
        return self.conv(_cuda_only(x, "WeightNormConv2d"))


def _wn(cin, cout, k, bias, weight_norm, scale):
    return WeightNormConv2d(cin, cout, (k, k), stride=1, padding=k // 2, bias=bias, weight_norm=weight_norm, scale=scale)


def _conv_bn_relu_chain(cin, dim, cout, bottleneck, weight_norm):
    """conv [BN ReLU conv]* as the reference stacks them: 1x1-3x3-1x1 with the bottleneck, 3x3-3x3 without; only
    the last conv has a bias and a trainable magnitude (modules_realnvp.py:86-105, 153-173)."""
    kernels = (1, 3, 1) if bottleneck else (3, 3)
    dims = [cin] + [dim] * (len(kernels) - 1) + [cout]
    layers = []
    for i, k in enumerate(kernels):
        last = i == len(kernels) - 1
        if i:
            layers += [nn.BatchNorm2d(dims[i]), nn.ReLU()]
        layers.append(_wn(dims[i], dims[i + 1], k, last, weight_norm, last))
    return nn.Sequential(*layers)


class ResidualBlock(nn.Module):
    """x + convs(ReLU(BN(x))) (modules_realnvp.py:73-114)."""

    def __init__(self, dim, bottleneck, weight_norm):
        super().__init__()
        self.in_block = nn.Sequential(nn.BatchNorm2d(dim), nn.ReLU())
        self.res_block = _conv_bn_relu_chain(dim, dim, dim, bottleneck, weight_norm)

    def forward(self, x):
        """Stand-alone call (ATen on the CUDA device); inside a native coupling the fused kernels evaluate it."""
        _cuda_only(x, "ResidualBlock")
        return x + self.res_block(self.in_block(x))


class ResidualModule(nn.Module):
    """The s/t network: in conv, ``res_blocks`` residual blocks with 1x1 skip taps, out conv; or, with
    ``res_blocks == 0``, one plain conv chain (modules_realnvp.py:116-194)."""

    def __init__(self, in_dim, dim, out_dim, res_blocks, bottleneck, skip, weight_norm):
        super().__init__()
        self.res_blocks, self.skip = res_blocks, skip
        # construction order = the reference's, so seeds reproduce its initialisation
        if res_blocks > 0:
            self.in_block = _wn(in_dim, dim, 3, True, weight_norm, False)
            self.core_block = nn.ModuleList(ResidualBlock(dim, bottleneck, weight_norm) for _ in range(res_blocks))
            self.out_block = nn.Sequential(nn.BatchNorm2d(dim), nn.ReLU(), _wn(dim, out_dim, 1, True, weight_norm, True))
            if skip:
                self.in_skip = _wn(dim, dim, 1, True, weight_norm, True)
                self.core_skips = nn.ModuleList(_wn(dim, dim, 1, True, weight_norm, True) for _ in range(res_blocks))
        else:
            self.block = _conv_bn_relu_chain(in_dim, dim, out_dim, bottleneck, weight_norm)

    def forward(self, x):
        """Stand-alone call (ATen on the CUDA device); inside a native coupling the fused kernels evaluate it."""
        _cuda_only(x, "ResidualModule")
        if self.res_blocks == 0:
            return self.block(x)
        a = self.in_block(x)
        taps = self.in_skip(a) if self.skip else None
        for i, blk in enumerate(self.core_block):
            a = blk(a)
            if self.skip:
                taps = taps + self.core_skips[i](a)
        return self.out_block(taps if self.skip else a)


class AbstractCoupling(nn.Module):
    """Shared bookkeeping of the two coupling kinds (modules_realnvp.py:196-237)."""

    _KIND = -1

    def __init__(self, mask_config, hps):
        super().__init__()
        self.mask_config = mask_config
        self.res_blocks = hps.res_blocks
        self.bottleneck = hps.bottleneck
        self.skip = hps.skip
        self.weight_norm = hps.weight_norm
        self.coupling_bn = hps.coupling_bn
        self._native = _native_hps(hps)
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
        for e in self._all_engines():
            e.dirty = True
        if "mask" in self.__dict__ and isinstance(self.mask, torch.Tensor):
            self.mask = fn(self.mask)      # the reference keeps it a plain attribute on the device (:251-254)
        return out

    def _shape(self):
        raise NotImplementedError

    def _all_engines(self):
        out = [self._engine] if self._engine is not None else []
        return out + list(self.__dict__.get("_engines_by_size", {}).values())

    def set_math(self, mode):
        """'tf32' / 'fp32' arithmetic tier for this module's stand-alone calls (None = default)."""
        import rnvp_cabi
        self._math = None if mode is None else rnvp_cabi.MATH_BY_NAME[mode]
        for e in self._all_engines():
            e.set_math(_eng._DEFAULT_MATH if self._math is None else self._math)

    def _engine_for(self, x):
        raise NotImplementedError

    def forward(self, x, reverse=False):
        """Returns ``(transformed x, log_diag_J)`` like the reference (modules_realnvp.py:264, 324)."""
        self._check_input(x)
        if not self._native:
            return self._aten_forward(_cuda_only(x, type(self).__name__), reverse)
        eng = self._engine_for(x)
        if reverse:
            with torch.no_grad():
                return eng.coupling_inverse(0, x, self.training)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _eng.CouplingFn.apply(_eng.grad_anchor(x.device), x, eng, self.training)
        y, logj, _ = eng.coupling_forward(0, x, self.training)
        return y, logj

    # -- the hyper-parameter branches train.py never selects: ATen operators on the device -------------------- #
    def _split(self, x):
        """(identity part fed to the s/t net, part to transform, keep-mask of the transform or None)."""
        raise NotImplementedError

    def _merge(self, x, on, logj):
        raise NotImplementedError

    def _aten_forward(self, x, reverse):
        cond, on, keep = self._split(x)
        u = self.in_bn(cond)
        h = torch.cat((u, -u) if keep is None else (u, -u, (1.0 - keep).expand(x.shape[0], -1, -1, -1)), dim=1)
        t, l = self.block(h).chunk(2, dim=1)
        s = self.scale * torch.tanh(l) + self.scale_shift
        if keep is not None:
            t, s = t * keep, s * keep
        k = 1.0 if keep is None else keep
        logj = s
        rm = self.out_bn.running_mean.view(1, -1, 1, 1)
        rv = self.out_bn.running_var.view(1, -1, 1, 1)
        if reverse:
            if self.coupling_bn:                               # always the RUNNING statistics (:285-291)
                on = on * torch.exp(0.5 * torch.log(rv + 1e-5) * k) + rm * k
            on = (on - t) * torch.exp(-s)
        else:
            on = on * torch.exp(s) + t
            if self.coupling_bn:
                var = self.batch_stat(on)[1] if self.training else rv
                normed = self.out_bn(on)
                on = normed if keep is None else normed * keep + on * (1.0 - keep)
                logj = s - 0.5 * torch.log(var + 1e-5) * k
        return self._merge(x, on, logj)

    def _check_input(self, x):
        c, s, _ = self._shape()
        if x.dim() != 4 or x.shape[1] != c or (s is not None and (x.shape[2] != s or x.shape[3] != s)):
            raise ValueError(f"expected input of shape (B,{c},{s},{s}), got {tuple(x.shape)}")


class CheckerboardAffineCoupling(AbstractCoupling):
    """Affine coupling with a checkerboard mask (modules_realnvp.py:239-302)."""

    _KIND = 0

    def __init__(self, in_out_dim, mid_dim, size, mask_config, hps):
        super().__init__(mask_config, hps)
        self.mask = self.build_mask(size, config=mask_config)     # kept for API parity; the kernels derive it
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

    def _engine_for(self, x):
        if self._engine is None:
            c, s, d = self._dims
            self._engine = _eng.Engine.for_coupling(0, c, s, d, int(self.mask_config), self.res_blocks, self,
                                                    math=self._math)
        return self._engine

    def _split(self, x):
        m = self.mask.to(x.device)
        return x * m, x, 1.0 - m

    def _merge(self, x, on, logj):
        return on, logj


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

    def _engine_for(self, x):
        # the spatial size is not a constructor argument of this class: plans are keyed by it
        s = x.shape[2]
        engines = self.__dict__.setdefault("_engines_by_size", {})
        if s not in engines:
            c, _, d = self._dims
            engines[s] = _eng.Engine.for_coupling(1, c, s, d, int(self.mask_config), self.res_blocks, self, math=self._math)
        return engines[s]

    def _split(self, x):
        a, b = x.chunk(2, dim=1)
        return (b, a, None) if self.mask_config else (a, b, None)

    def _merge(self, x, on, logj):
        cond = x.chunk(2, dim=1)[1 if self.mask_config else 0]
        zeros = torch.zeros_like(logj)
        if self.mask_config:
            return torch.cat((on, cond), dim=1), torch.cat((logj, zeros), dim=1)
        return torch.cat((cond, on), dim=1), torch.cat((zeros, logj), dim=1)
