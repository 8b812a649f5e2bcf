"""Host-side glue between the drop-in ``nn.Module`` classes and the C-ABI plan.

An :class:`Engine` owns one ``rnvp_plan`` (a whole multi-scale stack or a single
stand-alone coupling), the parameter / gradient pointer table handed to
``rnvp_plan_bind`` and the caller-owned workspace.  PyTorch stays the owner of
every tensor; the library only borrows raw device pointers for the duration of a
call (SURVEY.md 8b "Ownership").

Gradients: the backward entry points ADD parameter gradients into a flat fp32
buffer whose per-parameter views are installed as ``param.grad``.  One
``autograd.Function`` node covers the whole stack, so a training step costs a
handful of Python calls instead of ~10^4 autograd nodes, and the flat buffer is
what the data-parallel wrapper all-reduces in buckets.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

import rnvp_cabi as cabi
from rnvp_cabi import check, lib, ptr

_DEFAULT_MATH = cabi.MATH_TF32


def set_default_math(mode: str) -> None:
    """'fp32' (CUDA-core, 1e-5 tier) or 'tf32' (tcgen05 tensor cores, 1e-3 tier)."""
    global _DEFAULT_MATH
    _DEFAULT_MATH = cabi.MATH_BY_NAME[mode]


def _resolve(module: nn.Module, dotted: str):
    obj = module
    for part in dotted.split("."):
        obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
    return obj


def _require_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: this package has no CPU path "
                           "(the RealNVP hot path runs on sm_100a kernels only)")
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {t.dtype}")
    return t.contiguous()


class Engine:
    def __init__(self, handle: C.c_void_p, couplings: Sequence[nn.Module], math: Optional[int] = None):
        self.handle = handle
        self.couplings = list(couplings)
        self.n_cpl = lib.rnvp_plan_num_couplings(handle)
        assert self.n_cpl == len(self.couplings), (self.n_cpl, len(self.couplings))
        self.slots = lib.rnvp_plan_slots_per_coupling(handle)
        self.slot_names = [(lib.rnvp_plan_slot_name(handle, i) or b"").decode() for i in range(self.slots)]
        self.math = _DEFAULT_MATH if math is None else math
        check(lib.rnvp_plan_set_math(handle, self.math))
        self._bound_sig = None
        self._tensors: List[Optional[torch.Tensor]] = []
        self._trainable: List[nn.Parameter] = []
        self._flat_grad: Optional[torch.Tensor] = None
        self.grad_writes = 0                # backward passes that have added into the flat gradient (rnvp_optim)
        self._views: List[torch.Tensor] = []
        self._ws = {0: None, 1: None, 2: None}
        self.train_mode = None              # None = pick 2 (keep activations) when memory allows, else 1
        self.dirty = True
        self.dp = None                      # set by rnvp_dp.DataParallel

    # -- construction helpers ----------------------------------------------------- #
    @classmethod
    def for_flow(cls, channels, image_size, base_dim, res_blocks, num_scales, prior_loc, prior_scale,
                 couplings, math=None) -> "Engine":
        cfg = cabi.Config(channels, image_size, base_dim, res_blocks, num_scales, prior_loc, prior_scale)
        h = C.c_void_p()
        check(lib.rnvp_plan_create(C.byref(cfg), C.byref(h)))
        return cls(h, couplings, math)

    @classmethod
    def for_coupling(cls, kind, c, s, d, mask_cfg, res_blocks, module, math=None) -> "Engine":
        h = C.c_void_p()
        check(lib.rnvp_plan_create_single(kind, c, s, d, int(mask_cfg), res_blocks, C.byref(h)))
        return cls(h, [module], math)

    def __del__(self):
        try:
            if self.handle:
                lib.rnvp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_math(self, math: int) -> None:
        self.math = math
        check(lib.rnvp_plan_set_math(self.handle, math))

    # -- parameter table ---------------------------------------------------------- #
    def _collect(self):
        tensors: List[Optional[torch.Tensor]] = []
        for m in self.couplings:
            for name in self.slot_names:
                tensors.append(_resolve(m, name) if name else None)
        return tensors

    def bind(self, device: torch.device) -> None:
        tensors = self._collect()
        for t in tensors:
            if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError("all RealNVP parameters and buffers must be contiguous float32 CUDA tensors "
                                   "(call model.to('cuda') first); there is no CPU path")
        trainable = [t for t in tensors if isinstance(t, nn.Parameter) and t.requires_grad]
        total = sum(t.numel() for t in trainable)
        if self._flat_grad is None or self._flat_grad.numel() != total or self._flat_grad.device != device:
            self._flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        views, off = [], 0
        for t in trainable:
            views.append(self._flat_grad[off:off + t.numel()].view_as(t))
            off += t.numel()
        view_of = {id(t): v for t, v in zip(trainable, views)}
        n = len(tensors)
        params = (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in tensors])
        grads = (C.c_void_p * n)(*[view_of[id(t)].data_ptr() if (t is not None and id(t) in view_of) else None
                                   for t in tensors])
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib.rnvp_plan_bind(self.handle, params, grads, C.c_void_p(stream)))
        self._tensors, self._trainable, self._views = tensors, trainable, views
        # num_batches_tracked counters (nn.BatchNorm2d bumps them on every train-mode call)
        self._nbt_all, self._nbt_inv = [], []
        for m in self.couplings:
            for name, mod in m.named_modules():
                if isinstance(mod, nn.BatchNorm2d) and mod.num_batches_tracked is not None:
                    self._nbt_all.append(mod.num_batches_tracked)
                    if name != "out_bn":
                        self._nbt_inv.append(mod.num_batches_tracked)
        self._bound_sig = self._signature(tensors)
        self.dirty = False
        # per-coupling slices of the flat gradient, in forward order (DP buckets)
        self.cpl_grad_ranges = []
        off = 0
        per = self.slots
        for ci in range(self.n_cpl):
            cnt = sum(t.numel() for t in tensors[ci * per:(ci + 1) * per]
                      if isinstance(t, nn.Parameter) and t.requires_grad)
            self.cpl_grad_ranges.append((off, off + cnt))
            off += cnt

    @staticmethod
    def _signature(tensors):
        first = next(t for t in tensors if t is not None)
        last = next(t for t in reversed(tensors) if t is not None)
        return (first.data_ptr(), last.data_ptr(), len(tensors))

    def ensure_bound(self, device: torch.device) -> None:
        if self.dirty or self._bound_sig is None:
            self.bind(device)
            return
        t = self._tensors
        first = next(x for x in t if x is not None)
        last = next(x for x in reversed(t) if x is not None)
        if (first.data_ptr(), last.data_ptr(), len(t)) != self._bound_sig or \
                _resolve(self.couplings[0], "scale") is not first:
            self.bind(device)

    def pick_train_mode(self, batch: int, device: torch.device) -> int:
        if self.train_mode is not None:
            return self.train_mode
        need2 = lib.rnvp_plan_workspace_bytes(self.handle, batch, 2)
        ws = self._ws[2]
        if ws is not None and ws.numel() >= need2 and ws.device == device:
            return 2
        free, _total = torch.cuda.mem_get_info(device)
        held = sum(w.numel() for w in self._ws.values() if w is not None and w.device == device)
        return 2 if need2 < 0.8 * (free + held) else 1

    def workspace(self, batch: int, mode: int, device: torch.device) -> torch.Tensor:
        need = lib.rnvp_plan_workspace_bytes(self.handle, batch, mode)
        ws = self._ws[mode]
        if ws is None or ws.numel() < need or ws.device != device:
            self._ws[mode] = None
            if mode:                       # the two training layouts never coexist
                self._ws[3 - mode] = None
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[mode] = ws
        return ws

    def release_workspaces(self) -> None:
        """Drop the cached workspaces (e.g. the training layout before a large sampling batch)."""
        for k in self._ws:
            self._ws[k] = None

    # -- gradient installation ------------------------------------------------------ #
    def prepare_grads(self) -> None:
        """Make ``param.grad`` of every trainable parameter alias its view of the flat buffer.

        ``None`` grads (``optimizer.zero_grad()`` default) become zeroed views; a foreign gradient
        tensor is copied into its view first, so accumulation semantics are those of autograd.
        """
        tr, vs = self._trainable, self._views
        none_count, foreign = 0, []
        for i, p in enumerate(tr):
            g = p.grad
            if g is None:
                none_count += 1
            elif g.data_ptr() != vs[i].data_ptr():
                foreign.append(i)
        if none_count == len(tr):
            self._flat_grad.zero_()
            for p, v in zip(tr, vs):
                p.grad = v
            return
        if none_count == 0 and not foreign:
            return
        for i, p in enumerate(tr):
            if p.grad is None:
                vs[i].zero_()
                p.grad = vs[i]
        for i in foreign:
            vs[i].copy_(tr[i].grad)
            tr[i].grad = vs[i]

    # -- calls ------------------------------------------------------------------------ #
    def flow_forward(self, x: torch.Tensor, training: bool, want_z=False, want_ws=True):
        x = _require_cuda(x, "x")
        dev = x.device
        self.ensure_bound(dev)
        B = x.shape[0]
        mode = self.pick_train_mode(B, dev) if training else 0
        ws = self.workspace(B, mode, dev)
        ll = torch.empty(B, dtype=torch.float32, device=dev)
        logdet = torch.empty(B, dtype=torch.float32, device=dev)
        z = torch.empty_like(x) if want_z else None
        wsc = torch.empty((), dtype=torch.float32, device=dev) if want_ws else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_flow_forward(self.handle, ptr(x), ptr(ll), ptr(logdet), ptr(z), ptr(wsc), B,
                                    mode, ptr(ws), ws.numel(), C.c_void_p(stream)))
        if training:
            torch._foreach_add_(self._nbt_all, 1)
        return ll, logdet, z, wsc, ws

    def flow_backward(self, dll: torch.Tensor, dws: Optional[torch.Tensor], want_dx: bool, x_like: torch.Tensor,
                      ws: torch.Tensor):
        dev = dll.device
        B = dll.shape[0]
        self.prepare_grads()
        self.grad_writes += 1
        dx = torch.empty_like(x_like) if want_dx else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_flow_backward(self.handle, ptr(dll), ptr(dws), ptr(dx), B, ptr(ws), ws.numel(),
                                     C.c_void_p(stream)))
        if self.dp is not None:
            self.dp.reduce_gradients(self)
        return dx

    def flow_inverse(self, z: torch.Tensor, training: bool) -> torch.Tensor:
        z = _require_cuda(z, "z")
        dev = z.device
        self.ensure_bound(dev)
        B = z.shape[0]
        ws = self.workspace(B, 0, dev)
        x = torch.empty_like(z)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_flow_inverse(self.handle, ptr(z), ptr(x), B, 1 if training else 0, ptr(ws), ws.numel(),
                                    C.c_void_p(stream)))
        if training:
            torch._foreach_add_(self._nbt_inv, 1)
        return x

    def coupling_forward(self, idx: int, x: torch.Tensor, training: bool):
        x = _require_cuda(x, "x")
        dev = x.device
        self.ensure_bound(dev)
        B = x.shape[0]
        ws = self.workspace(B, 1 if training else 0, dev)
        y, logj = torch.empty_like(x), torch.empty_like(x)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_coupling_forward(self.handle, idx, ptr(x), ptr(y), ptr(logj), B, 1 if training else 0,
                                        ptr(ws), ws.numel(), C.c_void_p(stream)))
        if training:
            torch._foreach_add_(self._nbt_all, 1)
        return y, logj, ws

    def coupling_inverse(self, idx: int, y: torch.Tensor, training: bool):
        """(x, log_rescale) of ``coupling(y, reverse=True)`` (modules_realnvp.py:284-291, 345-351)."""
        y = _require_cuda(y, "x")
        dev = y.device
        self.ensure_bound(dev)
        B = y.shape[0]
        ws = self.workspace(B, 0, dev)
        x, logj = torch.empty_like(y), torch.empty_like(y)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_coupling_inverse(self.handle, idx, ptr(y), ptr(x), ptr(logj), B, 1 if training else 0, ptr(ws),
                                        ws.numel(), C.c_void_p(stream)))
        if training:
            torch._foreach_add_(self._nbt_inv, 1)
        return x, logj

    def coupling_backward(self, idx: int, dy: torch.Tensor, dlogj: torch.Tensor, ws: torch.Tensor) -> torch.Tensor:
        dev = dy.device
        B = dy.shape[0]
        self.prepare_grads()
        self.grad_writes += 1
        dx = torch.empty_like(dy)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.rnvp_coupling_backward(self.handle, idx, ptr(dy.contiguous()), ptr(dlogj.contiguous()), ptr(dx), B,
                                         ptr(ws), ws.numel(), C.c_void_p(stream)))
        return dx


# ----------------------------------------------------------------------------------- #
# autograd nodes                                                                      #
# ----------------------------------------------------------------------------------- #
class FlowLogProb(torch.autograd.Function):
    """(ll, weight_scale) = forward(x); one node for the whole multi-scale stack."""

    @staticmethod
    def forward(ctx, anchor, x, engine: Engine, training: bool):
        ll, _logdet, _z, wsc, ws = engine.flow_forward(x, training, want_z=False, want_ws=True)
        ctx.engine, ctx.ws, ctx.training = engine, ws, training
        ctx.gen = lib.rnvp_plan_forward_generation(engine.handle)
        ctx.x_meta = x
        ctx.want_dx = ctx.needs_input_grad[1]
        return ll, wsc

    @staticmethod
    def backward(ctx, dll, dws):
        if not ctx.training:
            raise RuntimeError("backward through RealNVP needs a train-mode forward (model.train()); "
                               "eval-mode forwards keep no activations")
        eng: Engine = ctx.engine
        _check_generation(eng, ctx.gen)
        if dll is None:
            dll = torch.zeros(ctx.x_meta.shape[0], dtype=torch.float32, device=ctx.x_meta.device)
        dll = dll.contiguous()
        dws = None if dws is None else dws.contiguous()
        dx = eng.flow_backward(dll, dws, ctx.want_dx, ctx.x_meta, ctx.ws)
        return None, dx, None, None


class CouplingFn(torch.autograd.Function):
    """(y, log_diag_J) = coupling(x) for a stand-alone coupling module."""

    @staticmethod
    def forward(ctx, anchor, x, engine: Engine, training: bool):
        y, logj, ws = engine.coupling_forward(0, x, training)
        ctx.engine, ctx.ws, ctx.training = engine, ws, training
        ctx.gen = lib.rnvp_plan_forward_generation(engine.handle)
        return y, logj

    @staticmethod
    def backward(ctx, dy, dlogj):
        if not ctx.training:
            raise RuntimeError("backward through a coupling needs a train-mode forward")
        _check_generation(ctx.engine, ctx.gen)
        if dy is None:
            dy = torch.zeros_like(dlogj)
        if dlogj is None:
            dlogj = torch.zeros_like(dy)
        dx = ctx.engine.coupling_backward(0, dy, dlogj, ctx.ws)
        return None, dx, None, None


def _check_generation(engine: Engine, gen: int) -> None:
    """The activations of a training forward live in the engine's (shared) workspace: a later forward / inverse
    call of the same module overwrites them, after which the earlier graph can no longer be back-propagated."""
    now = lib.rnvp_plan_forward_generation(engine.handle)
    if now != gen:
        raise RuntimeError(
            "backward of a stale RealNVP forward: the module ran another forward / inverse pass since this graph "
            f"was built (generation {gen} -> {now}) and its saved activations were overwritten; call backward() "
            "before the next forward of the same module")


def grad_anchor(device) -> torch.Tensor:
    """A scalar that requires grad so the custom nodes are always part of the graph."""
    return torch.zeros((), device=device, requires_grad=True)
