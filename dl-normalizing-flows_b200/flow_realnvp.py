"""Drop-in replacement for the reference's ``flow_realnvp.py`` (B200-native compute).

``RealNVP(channels, image_size, prior, hps)`` has the reference's constructor,
children (``s1_ckbd`` ... ``s5_ckbd``), public methods and ``state_dict``
(flow_realnvp.py:35-370).  ``forward`` / ``log_prob`` / ``sample``
/ ``g`` run the whole multi-scale stack through ONE call into the C-ABI
(``rnvp_flow_forward`` / ``rnvp_flow_inverse``), and the backward pass through
one ``rnvp_flow_backward``; there is no torch operator on those paths and no CPU
fallback.  ``f`` keeps the reference's ``(z, log_diag_J)`` contract with the full
Jacobian-diagonal tensor and therefore walks the couplings one by one.

``num_scales`` (keyword, default 5 = the in-tree model) generalises the
hard-coded five scales so that BASELINE config 3 (two scales) can be built.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

import rnvp_engine as _eng
from rnvp_cabi import check, lib, ptr
from modules_realnvp import ChannelwiseAffineCoupling, CheckerboardAffineCoupling, _native_hps

_FACTOR_TAPS = ((0, 0), (1, 1), (0, 1), (1, 0))        # k -> (dy, dx) of factor_out (flow_realnvp.py:148-164)


# ATen forms of the layout maps (differentiable): used by the hyper-parameter branches that run on torch operators
def _t_squeeze(x):
    B, Cc, H, W = x.shape
    return x.reshape(B, Cc, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, Cc * 4, H // 2, W // 2)


def _t_undo_squeeze(x):
    B, Cc, H, W = x.shape
    return x.reshape(B, Cc // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3).reshape(B, Cc // 4, H * 2, W * 2)


def _t_factor_out(x):
    full = torch.cat([x[:, :, dy::2, dx::2] for dy, dx in _FACTOR_TAPS], dim=1)
    return full.chunk(2, dim=1)


def _t_restore(on, off):
    full = torch.cat((on, off), dim=1)
    B, C4, H, W = full.shape
    out = full.new_zeros(B, C4 // 4, H * 2, W * 2)
    for k, (dy, dx) in enumerate(_FACTOR_TAPS):
        out[:, :, dy::2, dx::2] = full[:, k * (C4 // 4):(k + 1) * (C4 // 4)]
    return out


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class RealNVP(nn.Module):
    def __init__(self, channels, image_size, prior, hps, num_scales=5):
        super().__init__()
        self.prior = prior
        self.channels = channels
        self.image_size = image_size
        self.num_scales = num_scales
        self._hps = hps
        self._engine = None
        # the configuration train.py constructs runs on the sm_100a kernels; the other hyper-parameter branches
        #  compose the modules' ATen forwards on the device
        self._native = _native_hps(hps)
        if image_size % (1 << (num_scales - 1)):
            raise ValueError(f"image_size={image_size} must be divisible by 2^{num_scales - 1}")
        chan, size, dim = channels, image_size, hps.base_dim
        for s in range(1, num_scales):
            # registration order (ckbd, chan) per scale = the reference's (flow_realnvp.py:51-88)
            setattr(self, f"s{s}_ckbd", self.checkerboard_combo(chan, dim, size, hps))
            setattr(self, f"s{s}_chan", self.channelwise_combo(chan * 4, dim * 2, hps))
            setattr(self, f"order_matrix_{s}", self.order_matrix(chan))
            chan, size, dim = chan * 2, size // 2, dim * 2
        setattr(self, f"s{num_scales}_ckbd", self.checkerboard_combo(chan, dim, size, hps, final=True))

    # -- construction (flow_realnvp.py:98-116) ------------------------------------------ #
    def checkerboard_combo(self, in_out_dim, mid_dim, size, hps, final=False):
        configs = (1., 0., 1., 0.) if final else (1., 0., 1.)
        return nn.ModuleList(CheckerboardAffineCoupling(in_out_dim, mid_dim, size, c, hps) for c in configs)

    def channelwise_combo(self, in_out_dim, mid_dim, hps):
        return nn.ModuleList(ChannelwiseAffineCoupling(in_out_dim, mid_dim, c, hps) for c in (0., 1., 0.))

    def _groups(self):
        L = self.num_scales
        for s in range(1, L):
            yield s, "ckbd", getattr(self, f"s{s}_ckbd")
            yield s, "chan", getattr(self, f"s{s}_chan")
        yield L, "ckbd", getattr(self, f"s{L}_ckbd")

    def _couplings(self):
        return [m for _, _, grp in self._groups() for m in grp]

    # -- engine ------------------------------------------------------------------------- #
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine.dirty = True
        return out

    def engine(self) -> "_eng.Engine":
        if self._engine is None:
            loc = float(torch.as_tensor(self.prior.loc).reshape(-1)[0])
            scale = float(torch.as_tensor(self.prior.scale).reshape(-1)[0])
            h = self._hps
            self._engine = _eng.Engine.for_flow(self.channels, self.image_size, h.base_dim, h.res_blocks,
                                                self.num_scales, loc, scale, self._couplings())
        return self._engine

    def set_math(self, mode: str) -> None:
        """'tf32' (tcgen05 tensor cores, default) or 'fp32' (CUDA-core fp32, 1e-5 parity tier)."""
        import rnvp_cabi
        if not self._native:
            return
        self.engine().set_math(rnvp_cabi.MATH_BY_NAME[mode])
        for cpl in self._couplings():
            cpl.set_math(mode)

    # -- layout transforms (flow_realnvp.py:121-193) -------------------------------------- #
    def squeeze(self, x):
        x = _eng._require_cuda(x, "x")
        B, Cc, H, W = x.shape
        y = x.new_empty(B, Cc * 4, H // 2, W // 2)
        check(lib.rnvp_squeeze(ptr(x), ptr(y), B, Cc, H, W, _stream(x)))
        return y

    def undo_squeeze(self, x):
        x = _eng._require_cuda(x, "x")
        B, Cc, H, W = x.shape
        y = x.new_empty(B, Cc // 4, H * 2, W * 2)
        check(lib.rnvp_undo_squeeze(ptr(x), ptr(y), B, Cc, H, W, _stream(x)))
        return y

    def order_matrix(self, channel):
        """The reference's 0/1 re-ordering kernel, shape (4C, C, 2, 2) (flow_realnvp.py:139-165).

        Output channel k*C + c picks input pixel (dy,dx)_k of channel c with
        k -> (0,0), (1,1), (0,1), (1,0).  Kept for API parity: ``factor_out`` / ``restore`` apply
        this fixed index map directly instead of convolving with the matrix.
        """
        w = np.zeros((4 * channel, channel, 2, 2), dtype="float32")
        for k, (dy, dx) in enumerate(((0, 0), (1, 1), (0, 1), (1, 0))):
            for c in range(channel):
                w[k * channel + c, c, dy, dx] = 1.0
        return torch.from_numpy(w)

    def factor_out(self, x, order_matrix=None):
        x = _eng._require_cuda(x, "x")
        B, Cc, H, W = x.shape
        on = x.new_empty(B, 2 * Cc, H // 2, W // 2)
        off = torch.empty_like(on)
        check(lib.rnvp_factor_out(ptr(x), ptr(on), ptr(off), B, Cc, H, W, _stream(x)))
        return on, off

    def restore(self, on, off, order_matrix=None):
        on, off = _eng._require_cuda(on, "on"), _eng._require_cuda(off, "off")
        B, Cc, H, W = on.shape
        x = on.new_empty(B, Cc // 2, 2 * H, 2 * W)
        check(lib.rnvp_restore(ptr(on), ptr(off), ptr(x), B, Cc, H, W, _stream(on)))
        return x

    # -- z -> x and x -> z ------------------------------------------------------------------ #
    def g(self, z):
        """Inverse pass (flow_realnvp.py:196-249): one fused call, no autograd."""
        if not self._native:
            return self._aten_g(z)
        with torch.no_grad():
            return self.engine().flow_inverse(z, self.training)

    def _aten_g(self, z):
        x, offs = z, []
        for _ in range(1, self.num_scales):
            x, off = _t_factor_out(x)
            offs.append(off)
        for s, kind, group in reversed(list(self._groups())):
            if s < self.num_scales and kind == "chan":
                x = _t_squeeze(_t_restore(x, offs[s - 1]))
            for cpl in reversed(group):
                x, _ = cpl(x, reverse=True)
            if kind == "chan":
                x = _t_undo_squeeze(x)
        return x

    def f(self, x):
        """x -> (z, log_diag_J) with the full Jacobian-diagonal tensor (flow_realnvp.py:252-327).

        API-parity path: runs coupling by coupling so that log_diag_J can be carried through the
        squeeze / factor_out permutations like the reference does.  ``log_prob`` does not use it.
        """
        sq, usq, fo, rs = ((self.squeeze, self.undo_squeeze, self.factor_out, self.restore) if self._native else
                           (_t_squeeze, _t_undo_squeeze, _t_factor_out, _t_restore))
        z, J = x, torch.zeros_like(x)
        z_off, J_off = [], []
        for s, kind, group in self._groups():
            if kind == "chan":
                z, J = sq(z), sq(J)
            for cpl in group:
                z, inc = cpl(z)
                J = J + inc
            if kind == "chan":
                z, J = usq(z), usq(J)
                z, zo = fo(z)
                J, Jo = fo(J)
                z_off.append(zo)
                J_off.append(Jo)
        for zo, Jo in zip(reversed(z_off), reversed(J_off)):
            z, J = rs(z, zo), rs(J, Jo)
        return z, J

    def log_prob(self, x):
        """Per-sample log-likelihood (flow_realnvp.py:329-340), differentiable."""
        return self._log_prob_ws(x)[0]

    def _log_prob_ws(self, x):
        if not self._native:
            if not x.is_cuda:
                raise RuntimeError("x must be a CUDA tensor: this package has no CPU path")
            z, J = self.f(x)
            ll = J.sum(dim=(1, 2, 3)) + self.prior.log_prob(z).sum(dim=(1, 2, 3))
            ws = None
            for name, p in self.named_parameters():             # flow_realnvp.py:362-369
                if p.requires_grad and name.split(".")[-1] in ("weight_g", "scale"):
                    ws = p.pow(2).sum() if ws is None else ws + p.pow(2).sum()
            return ll, ws
        eng = self.engine()
        if torch.is_grad_enabled():
            return _eng.FlowLogProb.apply(_eng.grad_anchor(x.device), x, eng, self.training)
        ll, _ld, _z, wsc, _ = eng.flow_forward(x, self.training)
        return ll, wsc

    def latent(self, x):
        """(z, per-sample log-det, per-sample log-lik) from the fused path (no full J tensor)."""
        if not self._native:
            with torch.no_grad():
                z, J = self.f(x)
                ld = J.sum(dim=(1, 2, 3))
                return z, ld, ld + self.prior.log_prob(z).sum(dim=(1, 2, 3))
        with torch.no_grad():
            ll, ld, z, _w, _ = self.engine().flow_forward(x, self.training, want_z=True, want_ws=False)
        return z, ld, ll

    def sample(self, size):
        """prior.sample((size,C,H,W)) pushed through g (flow_realnvp.py:342-352)."""
        z = self.prior.sample((size, self.channels, self.image_size, self.image_size))
        return self.g(z)

    def forward(self, x):
        """(log-likelihood (B,), weight_scale) (flow_realnvp.py:354-370): weight_scale is the sum of
        squares of every trainable ``weight_g`` and ``scale``, computed by one reduction kernel."""
        return self._log_prob_ws(x)
