"""Fused Adam for the drop-in RealNVP.

The reference trains with ``torch.optim.Adam(model.parameters(), lr, weight_decay=...)`` (train.py:134,
``optimizer.step()`` at train.py:200).  That keeps working unchanged on the drop-in model.  This class
does the same update -- coupled L2 weight decay, bias-corrected moments, no amsgrad -- as ONE kernel
launch over the engine's flat gradient buffer (``rnvp_adam_step``) instead of a multi-tensor sweep over
1960 parameter tensors, and can clear the gradients in the same pass.

It is a ``torch.optim.Optimizer``: ``param_groups`` and ``state_dict()`` / ``load_state_dict()`` have
torch's layout (one group indexing all 2212 parameters, per-parameter ``step`` / ``exp_avg`` /
``exp_avg_sq`` for the trainable ones), so ``realnvp_state_optim.pt`` files written by either optimizer
load into the other (train.py:150, 250).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from rnvp_cabi import check, lib, ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 fused_zero_grad: bool = True):
        net = getattr(model, "module", model)            # accepts rnvp_dp.DataParallel
        if not hasattr(net, "engine"):
            raise TypeError("rnvp_optim.Adam drives a flow_realnvp.RealNVP (or its DataParallel wrapper)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= weight_decay or not all(0.0 <= b < 1.0 for b in betas):
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(net.parameters(), defaults)
        self._net = net
        self._fused_zero = bool(fused_zero_grad)
        self._handle: Optional[C.c_void_p] = None
        self._sig = None
        self._m = self._v = None
        self._step_t: Optional[torch.Tensor] = None
        self._clean_at = -1                # engine.grad_writes value right after a step that cleared the gradients

    # ------------------------------------------------------------------------------------------ #
    def _ensure(self):
        eng = self._net.engine()
        dev = next(self._net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("rnvp_optim.Adam needs the model on its CUDA device; there is no CPU path")
        eng.ensure_bound(dev)
        eng.prepare_grads()
        flat = eng._flat_grad
        sig = (flat.data_ptr(), flat.numel(), len(eng._trainable), eng._trainable[0].data_ptr())
        if sig == self._sig:
            return eng, flat
        if self._handle is not None:
            check(lib.rnvp_adam_destroy(self._handle))
            self._handle = None
        n = len(eng._trainable)
        base = flat.data_ptr()
        params = (C.c_void_p * n)(*[p.data_ptr() for p in eng._trainable])
        offs = (C.c_int64 * n)(*[(v.data_ptr() - base) // 4 for v in eng._views])
        sizes = (C.c_int64 * n)(*[p.numel() for p in eng._trainable])
        h = C.c_void_p()
        check(lib.rnvp_adam_create(params, offs, sizes, n, C.byref(h)))
        self._handle = h
        # moments share the flat gradient layout; carry over state that exists already (load_state_dict, rebind)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        step = None
        for p, gv in zip(eng._trainable, eng._views):
            o = (gv.data_ptr() - base) // 4
            st = self.state.get(p)
            mv, vv = m[o:o + p.numel()].view_as(p), v[o:o + p.numel()].view_as(p)
            if st:
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                step = float(st["step"]) if step is None else step
            self.state[p] = {"exp_avg": mv, "exp_avg_sq": vv}
        # one shared step counter object: updating 1960 scalars per step from Python is what this class avoids
        self._step_t = torch.tensor(0.0 if step is None else step, dtype=torch.float32)
        for p in eng._trainable:
            self.state[p]["step"] = self._step_t
        self._m, self._v, self._sig = m, v, sig
        return eng, flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng, flat = self._ensure()
        g = self.param_groups[0]
        self._step_t += 1
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        check(lib.rnvp_adam_step(self._handle, ptr(flat), ptr(self._m), ptr(self._v), flat.numel(), float(g["lr"]),
                                 float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                 float(g["weight_decay"]), int(self._step_t.item()), int(self._fused_zero),
                                 C.c_void_p(stream)))
        self._clean_at = eng.grad_writes if self._fused_zero else -1
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """Gradients live in one flat buffer whose views stay installed as ``param.grad``; clearing is a single
        memset, and nothing at all when the last ``step()`` already cleared them."""
        eng = self._net.engine()
        if eng._flat_grad is None or not eng._views:
            return super().zero_grad(set_to_none=set_to_none)
        if self._clean_at == eng.grad_writes:        # no backward has run since the clearing step
            return
        eng._flat_grad.zero_()

    def state_dict(self):
        """torch's layout; every entry gets its own ``step`` tensor and its own copy of the moments (internally
        the step counter is one shared object and the moments are views of two flat buffers, which torch's
        multi-tensor Adam must not inherit through ``load_state_dict``)."""
        sd = super().state_dict()
        sd["state"] = {k: {"step": v["step"].clone(), "exp_avg": v["exp_avg"].clone(),
                           "exp_avg_sq": v["exp_avg_sq"].clone()} for k, v in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._sig = None                   # re-flatten the loaded moments at the next step

    def __del__(self):
        try:
            if self._handle is not None:
                lib.rnvp_adam_destroy(self._handle)
        except Exception:
            pass
