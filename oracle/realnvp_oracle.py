"""CPU oracle for the RealNVP coupling-stack hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict driven) restatement, in plain torch CPU
ops, of the algorithm in the reference's ``flow_realnvp.py`` /
``modules_realnvp.py`` / ``utils.py``.  It exists so that the CUDA path can be
checked on a box where ``/root/reference`` is not mounted.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package never does.

Pinning status: the reference ships no tests / golden vectors (SURVEY.md §4,
§8c), so the oracle is pinned against *outputs of the reference itself*:
``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (in the build container), runs it and this oracle on the
same weights and inputs, asserts agreement, and commits the reference's
outputs as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py``
re-checks the oracle against those fixtures wherever the tests run.

All arithmetic the reference performs lives in the third-party dependency
``torch`` (un-pinned by the reference; torch 2.11.0 here), so this restatement
uses the same torch primitives (conv2d, batch statistics, tanh/exp/log).

Reference citations (file:line into /root/reference):
  weight norm ............ modules_realnvp.py:53-59  (old-style nn.utils.weight_norm, dim=0)
  ResidualBlock .......... modules_realnvp.py:73-114
  ResidualModule ......... modules_realnvp.py:116-194
  checkerboard mask ...... modules_realnvp.py:211-226
  batch_stat ............. modules_realnvp.py:228-237
  checkerboard coupling .. modules_realnvp.py:264-302
  channelwise coupling ... modules_realnvp.py:324-370
  squeeze / undo_squeeze . flow_realnvp.py:121-135
  order_matrix ........... flow_realnvp.py:139-165
  factor_out / restore ... flow_realnvp.py:167-193
  g (z -> x) ............. flow_realnvp.py:196-249
  f (x -> z) ............. flow_realnvp.py:252-327
  log_prob ............... flow_realnvp.py:329-340
  forward / weight_scale . flow_realnvp.py:354-370
  logit_transform ........ utils.py:33-72
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5          # nn.BatchNorm2d default, and the literal in modules_realnvp.py:289,301
BN_MOMENTUM = 0.1      # nn.BatchNorm2d default


# --------------------------------------------------------------------------- #
# TF32 operand emulation (tests of the tensor-core tier)                      #
# --------------------------------------------------------------------------- #
def round_tf32(x: Tensor) -> Tensor:
    """fp32 -> tf32 (10-bit mantissa) round-to-nearest, ties away from zero == PTX cvt.rna.tf32.f32."""
    if x.dtype != torch.float32:
        return x
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class _RoundOperand(torch.autograd.Function):
    """Forward: round to TF32.  Backward: straight through (the rounding is not differentiated)."""

    @staticmethod
    def forward(ctx, x):
        return round_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBoth(torch.autograd.Function):
    """A tensor that is a conv operand in BOTH passes: its value is rounded when it is produced and the
    gradient arriving at it (summed over all consumers) is rounded before it flows on."""

    @staticmethod
    def forward(ctx, x):
        return round_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return round_tf32(g)


class _RoundGrad(torch.autograd.Function):
    """Identity forward; the gradient w.r.t. this tensor is the ``dy`` operand of a dgrad / wgrad MMA."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return round_tf32(g)


# --------------------------------------------------------------------------- #
# topology                                                                    #
# --------------------------------------------------------------------------- #
def coupling_specs(channels: int, image_size: int, base_dim: int, num_scales: int = 5):
    """List of (name, kind, C, S, D, mask_config) in forward order.

    Follows flow_realnvp.py:40-95 (chan*=2, size//=2, dim*=2 per scale; the
    channelwise combo runs at 4*chan with mid_dim 2*dim; the last scale has four
    checkerboard couplings with configs 1,0,1,0) generalised to ``num_scales``.
    """
    specs = []
    chan, size, dim = channels, image_size, base_dim
    for s in range(1, num_scales):
        for i, cfg in enumerate((1, 0, 1)):                     # flow_realnvp.py:107-110
            specs.append((f"s{s}_ckbd.{i}", "ckbd", chan, size, dim, cfg))
        for i, cfg in enumerate((0, 1, 0)):                     # flow_realnvp.py:113-116
            specs.append((f"s{s}_chan.{i}", "chan", chan * 4, size // 2, dim * 2, cfg))
        chan, size, dim = chan * 2, size // 2, dim * 2
    for i, cfg in enumerate((1, 0, 1, 0)):                      # flow_realnvp.py:100-105
        specs.append((f"s{num_scales}_ckbd.{i}", "ckbd", chan, size, dim, cfg))
    return specs


def checkerboard_mask(size: int, config: int, dtype=torch.float32) -> Tensor:
    """mask[i,j] = (config + i + j) mod 2, shape (1,1,S,S) (modules_realnvp.py:223-226)."""
    m = np.arange(size).reshape(-1, 1) + np.arange(size)
    m = np.mod(config + m, 2).reshape(1, 1, size, size)
    return torch.tensor(m.astype("float64")).to(dtype)


# --------------------------------------------------------------------------- #
# layout transforms                                                           #
# --------------------------------------------------------------------------- #
def squeeze(x: Tensor) -> Tensor:
    """out[b,4c+2dy+dx,i,j] = in[b,c,2i+dy,2j+dx] (flow_realnvp.py:121-126)."""
    B, C, H, W = x.shape
    return x.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, C * 4, H // 2, W // 2)


def undo_squeeze(x: Tensor) -> Tensor:
    """Inverse of :func:`squeeze` (flow_realnvp.py:130-135)."""
    B, C, H, W = x.shape
    return x.reshape(B, C // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3).reshape(B, C // 4, H * 2, W * 2)


_FACTOR_TAPS = ((0, 0), (1, 1), (0, 1), (1, 0))   # k -> (dy,dx), decoded from flow_realnvp.py:148-164


def factor_out(x: Tensor) -> Tuple[Tensor, Tensor]:
    """full[b,k*C+c,i,j] = in[b,c,2i+dy_k,2j+dx_k]; on = k in {0,1}, off = k in {2,3}.

    Index-map form of the stride-2 0/1-kernel conv at flow_realnvp.py:177-180.
    """
    parts = [x[:, :, dy::2, dx::2] for (dy, dx) in _FACTOR_TAPS]
    full = torch.cat(parts, dim=1)
    C2 = full.shape[1] // 2
    return full[:, :C2], full[:, C2:]


def restore(on: Tensor, off: Tensor) -> Tensor:
    """Exact inverse of :func:`factor_out` (conv_transpose2d at flow_realnvp.py:192-193)."""
    full = torch.cat((on, off), dim=1)
    B, C4, H, W = full.shape
    C = C4 // 4
    out = full.new_zeros(B, C, H * 2, W * 2)
    for k, (dy, dx) in enumerate(_FACTOR_TAPS):
        out[:, :, dy::2, dx::2] = full[:, k * C:(k + 1) * C]
    return out


# --------------------------------------------------------------------------- #
# logit dequantisation (utils.py:33-72)                                       #
# --------------------------------------------------------------------------- #
def logit_forward(x: Tensor, noise: Tensor, constraint: float = 0.9) -> Tuple[Tensor, Tensor]:
    """x in [0,1] (ToTensor output), noise ~ U[0,1) supplied by the caller.

    Same op order as utils.py:49-72; the reference draws ``noise`` itself at
    utils.py:47, here it is an argument so both sides can share it.
    """
    x = (x * 255.0 + noise) / 256.0
    x = x * 2.0
    x = x - 1.0
    x = x * constraint
    x = x + 1.0
    x = x / 2.0
    y = torch.log(x) - torch.log(1.0 - x)
    pre = torch.tensor(np.log(constraint) - np.log(1.0 - constraint))
    ldj = F.softplus(y) + F.softplus(-y) - F.softplus(-pre)
    return y, torch.sum(ldj, dim=(1, 2, 3))


def logit_inverse(y: Tensor, constraint: float = 0.9) -> Tensor:
    """utils.py:34-42."""
    x = 1.0 / (torch.exp(-y) + 1.0)
    x = x * 2.0
    x = x - 1.0
    x = x / constraint
    x = x + 1.0
    x = x / 2.0
    return x


# --------------------------------------------------------------------------- #
# the model                                                                   #
# --------------------------------------------------------------------------- #
class RealNVPOracle:
    """Functional RealNVP over a reference-layout ``state_dict``.

    ``state`` maps the reference's state-dict keys to tensors; tensors that
    require grad take part in autograd, which is how tests obtain reference
    gradients.  Running statistics in ``state`` are updated in place by
    training-mode calls, like ``nn.BatchNorm2d`` does.
    """

    def __init__(self, state: Dict[str, Tensor], channels: int, image_size: int,
                 base_dim: int, res_blocks: int, num_scales: int = 5,
                 prior_loc: float = 0.0, prior_scale: float = 1.0, emulate_tf32: bool = False):
        # emulate_tf32: restate the arithmetic of the library's tensor-core tier -- every operand of a conv
        # MMA (activation, weight, and in the backward pass the output gradient) is rounded to TF32 by the
        # kernel that PRODUCES it (cvt.rna), products are exact, accumulation is fp32; everything that is
        # not a conv operand stays fp32.  Rounding points (dl-normalizing-flows_b200/csrc):
        #   forward : w; h0; every relu(bn(.)); the trunk a_i (conv epilogue)
        #   backward: d st; d skip-sum; d a_i; d u1, d u2 (BN-backward apply / dgrad epilogue outputs)
        self.emulate_tf32 = emulate_tf32
        self.state = state
        self.channels, self.image_size = channels, image_size
        self.base_dim, self.res_blocks, self.num_scales = base_dim, res_blocks, num_scales
        self.prior_loc, self.prior_scale = prior_loc, prior_scale
        self.specs = coupling_specs(channels, image_size, base_dim, num_scales)
        self.training = True
        self.update_running = True
        self.trace: Optional[Dict[str, Tensor]] = None     # filled with intermediates when a dict

    # -- small helpers ----------------------------------------------------- #
    def _p(self, key: str) -> Tensor:
        return self.state[key]

    def _rec(self, key: str, t: Tensor) -> Tensor:
        if self.trace is not None:
            if t.requires_grad:
                t.retain_grad()
            self.trace[key] = t
        return t

    def _wn_conv(self, prefix: str, x: Tensor, pad: int) -> Tensor:
        """w = g * v / ||v||_(1,2,3), then stride-1 conv (modules_realnvp.py:53-71)."""
        v, g = self._p(prefix + ".conv.weight_v"), self._p(prefix + ".conv.weight_g")
        w = v * (g / torch.linalg.vector_norm(v, dim=(1, 2, 3), keepdim=True))
        b = self.state.get(prefix + ".conv.bias")
        if self.emulate_tf32:
            w = _RoundOperand.apply(w)
        return F.conv2d(x, w, b, stride=1, padding=pad)

    # roles of a tensor under TF32 emulation (identity otherwise)
    def _act(self, h: Tensor) -> Tensor:          # produced only to be a conv input
        return _RoundOperand.apply(h) if self.emulate_tf32 else h

    def _trunk(self, a: Tensor) -> Tensor:        # conv input AND carrier of a dy operand
        return _RoundBoth.apply(a) if self.emulate_tf32 else a

    def _dy(self, t: Tensor) -> Tensor:           # conv output whose gradient is a dy operand
        return _RoundGrad.apply(t) if self.emulate_tf32 else t

    def _bn(self, prefix: str, x: Tensor, affine: bool = True, training: Optional[bool] = None) -> Tensor:
        """nn.BatchNorm2d semantics: batch stats (biased var) in training, running stats in eval."""
        training = self.training if training is None else training
        w = self._p(prefix + ".weight") if affine else None
        b = self._p(prefix + ".bias") if affine else None
        rm, rv = self._p(prefix + ".running_mean"), self._p(prefix + ".running_var")
        if training:
            if self.update_running:
                nbt = self.state.get(prefix + ".num_batches_tracked")
                if nbt is not None:
                    nbt += 1
                return F.batch_norm(x, rm, rv, w, b, True, BN_MOMENTUM, BN_EPS)
            return F.batch_norm(x, None, None, w, b, True, BN_MOMENTUM, BN_EPS)
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)

    # -- s/t network ------------------------------------------------------- #
    def _res_block(self, prefix: str, x: Tensor) -> Tensor:
        """Bottleneck block (modules_realnvp.py:83-97,114)."""
        h = self._act(F.relu(self._bn(prefix + ".in_block.0", x)))
        h = self._rec(prefix + ".u1", self._dy(self._wn_conv(prefix + ".res_block.0", h, 0)))
        h = self._act(F.relu(self._bn(prefix + ".res_block.1", h)))
        h = self._rec(prefix + ".u2", self._dy(self._wn_conv(prefix + ".res_block.3", h, 1)))
        h = self._act(F.relu(self._bn(prefix + ".res_block.4", h)))
        h = self._wn_conv(prefix + ".res_block.6", h, 0)
        return self._trunk(x + h)

    def _res_module(self, prefix: str, x: Tensor) -> Tensor:
        """ResidualModule with skip connections (modules_realnvp.py:175-194)."""
        a = self._rec(prefix + ".a0", self._trunk(self._wn_conv(prefix + ".in_block", self._act(x), 1)))
        out = self._wn_conv(prefix + ".in_skip", a, 0)
        for i in range(self.res_blocks):
            a = self._rec(f"{prefix}.a{i + 1}", self._res_block(f"{prefix}.core_block.{i}", a))
            out = out + self._wn_conv(f"{prefix}.core_skips.{i}", a, 0)
        out = self._dy(out)
        self._rec(prefix + ".skipsum", out)
        h = self._act(F.relu(self._bn(prefix + ".out_block.0", out)))
        return self._rec(prefix + ".st", self._dy(self._wn_conv(prefix + ".out_block.2", h, 0)))

    # -- couplings --------------------------------------------------------- #
    def coupling(self, name: str, x: Tensor, reverse: bool = False,
                 kind: Optional[str] = None, cfg: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """Run one coupling; ``kind``/``cfg`` default to the model's spec for ``name``."""
        if kind is None:
            spec = next(s for s in self.specs if s[0] == name)
            _, kind, _C, _S, _D, cfg = spec
        if kind == "ckbd":
            return self._ckbd(name, x, cfg, reverse)
        return self._chan(name, x, cfg, reverse)

    def _ckbd(self, name: str, x: Tensor, cfg: int, reverse: bool) -> Tuple[Tensor, Tensor]:
        """CheckerboardAffineCoupling.forward (modules_realnvp.py:264-302)."""
        B, C, S, _ = x.shape
        m = checkerboard_mask(S, cfg, x.dtype).repeat(B, 1, 1, 1)
        u = self._bn(name + ".in_bn", x * m)
        h = torch.cat((u, -u), dim=1)
        h = torch.cat((h, m), dim=1)
        h = F.relu(h)                                                   # block[0], :259
        st = self._res_module(name + ".block.1", h)
        shift, lr = st.split(C, dim=1)
        lr = self._p(name + ".scale") * torch.tanh(lr) + self._p(name + ".scale_shift")
        shift = shift * (1.0 - m)
        lr = lr * (1.0 - m)
        logJ = lr
        rm, rv = self._p(name + ".out_bn.running_mean"), self._p(name + ".out_bn.running_var")
        if reverse:
            mean = rm.reshape(1, -1, 1, 1)
            var = rv.reshape(1, -1, 1, 1)
            x = x * torch.exp(0.5 * torch.log(var + 1e-5) * (1.0 - m)) + mean * (1.0 - m)
            x = (x - shift) * torch.exp(-lr)
        else:
            x = x * torch.exp(lr) + shift
            self._rec(name + ".xprime", x)
            if self.training:
                mu = torch.mean(x, dim=(0, 2, 3), keepdim=True)
                var = torch.mean((x - mu) ** 2, dim=(0, 2, 3), keepdim=True)
            else:
                var = rv.reshape(1, -1, 1, 1)
            x = self._bn(name + ".out_bn", x, affine=False) * (1.0 - m) + x * m
            logJ = logJ - 0.5 * torch.log(var + 1e-5) * (1.0 - m)
        return x, logJ

    def _chan(self, name: str, x: Tensor, cfg: int, reverse: bool) -> Tuple[Tensor, Tensor]:
        """ChannelwiseAffineCoupling.forward (modules_realnvp.py:324-370)."""
        C = x.shape[1]
        if cfg:
            on, off = x.split(C // 2, dim=1)
        else:
            off, on = x.split(C // 2, dim=1)
        u = self._bn(name + ".in_bn", off)
        h = F.relu(torch.cat((u, -u), dim=1))
        st = self._res_module(name + ".block.1", h)
        shift, lr = st.split(C // 2, dim=1)
        lr = self._p(name + ".scale") * torch.tanh(lr) + self._p(name + ".scale_shift")
        logJ = lr
        rm, rv = self._p(name + ".out_bn.running_mean"), self._p(name + ".out_bn.running_var")
        if reverse:
            mean = rm.reshape(1, -1, 1, 1)
            var = rv.reshape(1, -1, 1, 1)
            on = on * torch.exp(0.5 * torch.log(var + 1e-5)) + mean
            on = (on - shift) * torch.exp(-lr)
        else:
            on = on * torch.exp(lr) + shift
            self._rec(name + ".xprime", on)
            if self.training:
                mu = torch.mean(on, dim=(0, 2, 3), keepdim=True)
                var = torch.mean((on - mu) ** 2, dim=(0, 2, 3), keepdim=True)
            else:
                var = rv.reshape(1, -1, 1, 1)
            on = self._bn(name + ".out_bn", on, affine=False)
            logJ = logJ - 0.5 * torch.log(var + 1e-5)
        if cfg:
            return torch.cat((on, off), dim=1), torch.cat((logJ, torch.zeros_like(logJ)), dim=1)
        return torch.cat((off, on), dim=1), torch.cat((torch.zeros_like(logJ), logJ), dim=1)

    # -- the multi-scale stack ------------------------------------------------ #
    def _group(self, s: int, kind: str) -> List[str]:
        return [n for (n, *_r) in self.specs if n.startswith(f"s{s}_{kind}.")]

    def f(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """x -> (z, log_diag_J) (flow_realnvp.py:252-327)."""
        z, J = x, torch.zeros_like(x)
        z_offs, J_offs = [], []
        for s in range(1, self.num_scales):
            for n in self._group(s, "ckbd"):
                z, inc = self.coupling(n, z)
                J = J + inc
            z, J = squeeze(z), squeeze(J)
            for n in self._group(s, "chan"):
                z, inc = self.coupling(n, z)
                J = J + inc
            z, J = undo_squeeze(z), undo_squeeze(J)
            z, zo = factor_out(z)
            J, Jo = factor_out(J)
            z_offs.append(zo)
            J_offs.append(Jo)
        for n in self._group(self.num_scales, "ckbd"):
            z, inc = self.coupling(n, z)
            J = J + inc
        for zo, Jo in zip(reversed(z_offs), reversed(J_offs)):
            z, J = restore(z, zo), restore(J, Jo)
        return z, J

    def g(self, z: Tensor) -> Tensor:
        """z -> x (flow_realnvp.py:196-249)."""
        x, offs = z, []
        for _ in range(1, self.num_scales):
            x, off = factor_out(x)
            offs.append(off)
        for n in reversed(self._group(self.num_scales, "ckbd")):
            x, _ = self.coupling(n, x, reverse=True)
        for s in range(self.num_scales - 1, 0, -1):
            x = restore(x, offs[s - 1])
            x = squeeze(x)
            for n in reversed(self._group(s, "chan")):
                x, _ = self.coupling(n, x, reverse=True)
            x = undo_squeeze(x)
            for n in reversed(self._group(s, "ckbd")):
                x, _ = self.coupling(n, x, reverse=True)
        return x

    def log_prob_parts(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """(z, log_det_J (B,), log_prior (B,)) -- flow_realnvp.py:337-339."""
        z, J = self.f(x)
        log_det = torch.sum(J, dim=(1, 2, 3))
        lp = (-((z - self.prior_loc) ** 2) / (2 * self.prior_scale ** 2)
              - math.log(self.prior_scale) - math.log(math.sqrt(2 * math.pi)))   # Normal.log_prob
        return z, log_det, torch.sum(lp, dim=(1, 2, 3))

    def log_prob(self, x: Tensor) -> Tensor:
        _, ld, lp = self.log_prob_parts(x)
        return lp + ld

    def weight_scale(self) -> Tensor:
        """Sum of p^2 over trainable ``*.weight_g`` and ``*.scale`` (flow_realnvp.py:362-369).

        Trainable weight_g are those of convs built with ``scale=True``:
        res_block.6, out_block.2, in_skip, core_skips.* (modules_realnvp.py:96-97,142-151).
        """
        total = None
        for k, p in self.state.items():
            last = k.split(".")[-1]
            if last == "scale" or (last == "weight_g" and is_trainable_g(k)):
                t = torch.pow(p, 2).sum()
                total = t if total is None else total + t
        return total

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        ws = self.weight_scale()
        return self.log_prob(x), ws


def is_trainable_g(key: str) -> bool:
    """True for weight_g of convs constructed with scale=True in the reference."""
    return (".res_block.6." in key or ".out_block.2." in key
            or ".in_skip." in key or ".core_skips." in key)


def is_trainable(key: str) -> bool:
    """Whether a state-dict key names a parameter with requires_grad=True in the reference."""
    last = key.split(".")[-1]
    if last in ("running_mean", "running_var", "num_batches_tracked"):
        return False
    if last == "weight_g":
        return is_trainable_g(key)
    return True


# --------------------------------------------------------------------------- #
# state construction (shapes follow SURVEY.md §8b; values are the caller's)   #
# --------------------------------------------------------------------------- #
def state_shapes(channels: int, image_size: int, base_dim: int, res_blocks: int,
                 num_scales: int = 5) -> Dict[str, Tuple[int, ...]]:
    """Every state-dict key of the reference model with its shape, in state-dict order."""
    out: Dict[str, Tuple[int, ...]] = {}
    for (name, kind, C, S, D, cfg) in coupling_specs(channels, image_size, base_dim, num_scales):
        coupling_state_shapes(name, kind, C, D, res_blocks, out)
    return out


def coupling_state_shapes(name: str, kind: str, C: int, D: int, res_blocks: int,
                          out: Optional[Dict[str, Tuple[int, ...]]] = None) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys/shapes of ONE coupling module, prefixed with ``name`` ('' for none)."""
    out = {} if out is None else out
    name = name + "." if name and not name.endswith(".") else name

    def bn(prefix, c, affine=True):
        if affine:
            out[prefix + ".weight"] = (c,)
            out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,)
        out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    def conv(prefix, cin, cout, k, bias):
        if bias:
            out[prefix + ".conv.bias"] = (cout,)
        out[prefix + ".conv.weight_g"] = (cout, 1, 1, 1)
        out[prefix + ".conv.weight_v"] = (cout, cin, k, k)

    if True:
        cin = 2 * C + 1 if kind == "ckbd" else C
        cio = C if kind == "ckbd" else C // 2
        out[name + "scale"] = (1,)
        out[name + "scale_shift"] = (1,)
        bn(name + "in_bn", cio)
        p = name + "block.1"
        conv(p + ".in_block", cin, D, 3, True)
        for i in range(res_blocks):
            q = f"{p}.core_block.{i}"
            bn(q + ".in_block.0", D)
            conv(q + ".res_block.0", D, D, 1, False)
            bn(q + ".res_block.1", D)
            conv(q + ".res_block.3", D, D, 3, False)
            bn(q + ".res_block.4", D)
            conv(q + ".res_block.6", D, D, 1, True)
        bn(p + ".out_block.0", D)
        conv(p + ".out_block.2", D, 2 * cio, 1, True)
        conv(p + ".in_skip", D, D, 1, True)
        for i in range(res_blocks):
            conv(f"{p}.core_skips.{i}", D, D, 1, True)
        bn(name + "out_bn", cio, affine=False)
    return out


def random_state(channels: int, image_size: int, base_dim: int, res_blocks: int,
                 num_scales: int = 5, seed: int = 0, dtype=torch.float32,
                 scale: float = 0.7, exercise: bool = True) -> Dict[str, Tensor]:
    """A well-conditioned random state for tests (NOT the reference's init).

    ``exercise`` randomises scale / scale_shift / BN affine / running stats so
    that tanh, exp and the log-det terms are all active.
    """
    return random_state_from_shapes(state_shapes(channels, image_size, base_dim, res_blocks, num_scales),
                                    seed=seed, dtype=dtype, scale=scale, exercise=exercise)


def random_state_from_shapes(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.float32,
                             scale: float = 0.7, exercise: bool = True) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    st: Dict[str, Tensor] = {}
    for k, shp in shapes.items():
        last = k.split(".")[-1]
        if last == "num_batches_tracked":
            st[k] = torch.zeros((), dtype=torch.int64)
        elif last == "weight_v":
            fan_in = shp[1] * shp[2] * shp[3]
            st[k] = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1).to(dtype) / math.sqrt(fan_in)
        elif last == "weight_g":
            if is_trainable_g(k):
                st[k] = (0.5 + torch.rand(shp, generator=g, dtype=torch.float64)).to(dtype) * 0.6
            else:
                st[k] = torch.ones(shp, dtype=dtype)
        elif last == "bias" and ".conv." in k:
            st[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * 0.1).to(dtype)
        elif last == "weight":                    # BN gamma
            st[k] = (1.0 + 0.2 * (torch.rand(shp, generator=g, dtype=torch.float64) - 0.5)).to(dtype) if exercise \
                else torch.ones(shp, dtype=dtype)
        elif last == "bias":                      # BN beta
            st[k] = (0.2 * (torch.rand(shp, generator=g, dtype=torch.float64) - 0.5)).to(dtype) if exercise \
                else torch.zeros(shp, dtype=dtype)
        elif last == "running_mean":
            st[k] = (0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype) if exercise \
                else torch.zeros(shp, dtype=dtype)
        elif last == "running_var":
            st[k] = (0.5 + torch.rand(shp, generator=g, dtype=torch.float64)).to(dtype) if exercise \
                else torch.ones(shp, dtype=dtype)
        elif last == "scale":
            st[k] = torch.full(shp, scale if exercise else 0.0, dtype=dtype)
        elif last == "scale_shift":
            st[k] = (0.05 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype) if exercise \
                else torch.zeros(shp, dtype=dtype)
        else:
            raise KeyError(k)
    return st


def synthetic_images(batch: int, channels: int, size: int, seed: int = 0) -> Tensor:
    """SURVEY.md §8d synthetic input: uint8 noise images as ``ToTensor`` would yield them."""
    g = torch.Generator().manual_seed(seed)
    x8 = torch.randint(0, 256, (batch, channels, size, size), generator=g, dtype=torch.uint8)
    return x8.float() / 255.0
