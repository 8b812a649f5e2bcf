"""Drop-in replacement for the RealNVP part of the reference's ``utils.py``.

``logit_transform`` (utils.py:33-72) runs as one coalesced sm_100a kernel: uniform
dequantisation noise is drawn in-kernel (Philox4x32-10 keyed from torch's CPU
generator, so ``torch.manual_seed`` still controls it), the [0.05, 0.95] squeeze,
the logit and the per-sample log-det reduction are fused.  The reference runs
this on the CPU before the host-to-device copy (train.py:187-189); here a CPU
input is moved to the current CUDA device first and the results stay there, so
the caller's ``.to(device)`` becomes a no-op.  uint8 images are accepted as well
(value/255 is what ``ToTensor`` yields), which shrinks the copy 4x.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from rnvp_cabi import check, lib, ptr


def _to_device(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        return x
    if not torch.cuda.is_available():
        raise RuntimeError("logit_transform: no CUDA device; rnvp-b200 has no CPU path")
    return x.to("cuda", non_blocking=True)


def logit_transform(x, constraint=0.9, reverse=False, noise=None):
    """Same contract as the reference: forward returns ``(logit_x, per-sample log-det)``,
    ``reverse=True`` returns ``(x, 0)``.  ``noise`` (optional, same shape, U[0,1)) replaces the
    in-kernel draw -- used by the parity tests to share the noise with the reference."""
    x = _to_device(x)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    if reverse:
        x = x.contiguous().float()
        out = torch.empty_like(x)
        check(lib.rnvp_logit_inverse(ptr(x), ptr(out), x.numel(), float(constraint), stream))
        return out, 0
    B = x.shape[0]
    n = x[0].numel() if B else 0
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    logdet = torch.empty(B, dtype=torch.float32, device=x.device)
    if noise is not None:
        noise = _to_device(noise).contiguous().float()
        seed = 0
    else:
        seed = int(torch.empty((), dtype=torch.int64).random_())      # CPU generator: follows manual_seed
    if x.dtype == torch.uint8:
        x = x.contiguous()
        check(lib.rnvp_logit_forward_u8(ptr(x), ptr(noise), ptr(y), ptr(logdet), B, n, float(constraint),
                                        seed & (2 ** 64 - 1), 0, stream))
    else:
        x = x.contiguous().float()
        check(lib.rnvp_logit_forward(ptr(x), ptr(noise), ptr(y), ptr(logdet), B, n, float(constraint),
                                     seed & (2 ** 64 - 1), 0, stream))
    return y, logdet


class Hyperparameters():
    """The hyper-parameter bag read by RealNVP and the coupling modules (utils.py:78-93)."""

    def __init__(self, base_dim, res_blocks, bottleneck, skip, weight_norm, coupling_bn):
        self.base_dim = base_dim
        self.res_blocks = res_blocks
        self.bottleneck = bottleneck
        self.skip = skip
        self.weight_norm = weight_norm
        self.coupling_bn = coupling_bn


def weights_init(m):
    """DCGAN initialiser imported by the reference's train.py (utils.py:98-113).  Not on the
    RealNVP path; provided so that ``from utils import weights_init`` keeps working."""
    name = type(m).__name__
    if "Conv" in name:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm" in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
