"""Launches the HBM-bound tensor-core kernels of the S = 64 / 32 scales a few times each, for `ncu --set full`
(GPU box only).  Order of the launches (NREP each, rotating over operand sets larger than L2):
  fwd_bn S64, fwd_bn S32 (BN-prologue 1x1 conv with bias + residual + statistics = rb6),
  wgrad S64, wgrad S32, wgrad_bn S64, wgrad_bn S32, dgrad_bn S32, fwd3x3 S32
Without ncu it prints the event-timed microseconds per launch of every case (pipelined, PDL active)."""
import importlib
import os
import sys
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch

pkg = importlib.import_module("dl-normalizing-flows_b200")
lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
import ctypes as C
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
B = int(os.environ.get("B", "256"))
NREP = int(os.environ.get("NREP", "3"))
TIME = os.environ.get("TIME", "0") == "1"
DEV = "cuda"


def make(S, D, k):
    P = B * S * S
    nset = 4
    d = dict(S=S, D=D, k=k, P=P, nset=nset)
    d["x"] = [torch.randn(B, S, S, D, device=DEV) for _ in range(nset)]
    d["y"] = [torch.empty(B, S, S, D, device=DEV) for _ in range(nset)]
    d["r"] = [torch.randn(B, S, S, D, device=DEV) for _ in range(nset)]
    d["w"] = torch.randn(k * k, D, D, device=DEV) * 0.05
    d["dw"] = torch.zeros(k * k, D, D, device=DEV)
    d["bias"] = torch.randn(D, device=DEV)
    d["db"] = torch.zeros(D, device=DEV)
    d["stats"] = torch.zeros(2 * D, dtype=torch.float64, device=DEV)
    d["sums"] = torch.cat((torch.zeros(D, dtype=torch.float64), torch.full((D,), float(P), dtype=torch.float64))).to(DEV)
    d["gamma"] = torch.ones(D, device=DEV)
    d["beta"] = torch.zeros(D, device=DEV)
    d["rm"] = torch.zeros(D, device=DEV)
    d["rv"] = torch.ones(D, device=DEV)
    d["save"] = torch.zeros(4 * D, device=DEV)
    d["save"][D:3 * D] = 1.0            # mean 0, rstd 1, scale 1, shift 0
    d["sums2"] = torch.zeros(2 * D, dtype=torch.float64, device=DEV)
    return d


def fwd_bn(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_forward_bn(ptr(d["x"][j]), ptr(d["w"]), ptr(d["bias"]), ptr(d["r"][j]), ptr(d["y"][j]), ptr(d["stats"]),
                                   B, d["S"], d["D"], d["D"], d["D"], d["k"], d["D"], 1, d["D"], ptr(d["sums"]), float(d["P"]),
                                   ptr(d["gamma"]), ptr(d["beta"]), ptr(d["rm"]), ptr(d["rv"]), ptr(d["save"]), 1, st))


def fwd(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_forward(ptr(d["x"][j]), ptr(d["w"]), None, None, ptr(d["y"][j]), ptr(d["stats"]), B, d["S"], d["D"],
                                d["D"], d["D"], d["k"], d["D"], 1, st))


def wgrad(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_wgrad(ptr(d["x"][j]), ptr(d["r"][j]), ptr(d["dw"]), None, B, d["S"], d["D"], d["D"], d["D"], d["k"],
                              d["D"], 1, st))


def wgrad_bn(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_wgrad_bn(ptr(d["x"][j]), ptr(d["r"][j]), ptr(d["dw"]), ptr(d["db"]), B, d["S"], d["D"], d["D"], d["D"],
                                 d["k"], d["D"], ptr(d["save"]), d["D"], st))


def wgrad_bias(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_wgrad(ptr(d["x"][j]), ptr(d["r"][j]), ptr(d["dw"]), ptr(d["db"]), B, d["S"], d["D"], d["D"], d["D"], d["k"],
                              d["D"], 1, st))


def wgrad_bn_nob(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_wgrad_bn(ptr(d["x"][j]), ptr(d["r"][j]), ptr(d["dw"]), None, B, d["S"], d["D"], d["D"], d["D"],
                                 d["k"], d["D"], ptr(d["save"]), d["D"], st))


def dgrad_bn(d, i):
    j = i % d["nset"]
    check(lib.rnvp_conv_dgrad_bn(ptr(d["r"][j]), ptr(d["w"]), ptr(d["x"][j]), ptr(d["save"]), ptr(d["y"][j]), ptr(d["sums2"]),
                                 B, d["S"], d["D"], d["D"], d["D"], d["k"], d["D"], st))


s64, s32 = make(64, 32, 1), make(32, 64, 1)
s32k3 = make(32, 64, 3)
CASES = [("fwd_bn S64 1x1", fwd_bn, s64), ("fwd_bn S32 1x1", fwd_bn, s32), ("wgrad S64 1x1", wgrad, s64),
         ("wgrad S32 1x1", wgrad, s32), ("wgrad_bn S64 1x1", wgrad_bn, s64), ("wgrad_bn S32 1x1", wgrad_bn, s32),
         ("dgrad_bn S32 1x1", dgrad_bn, s32), ("fwd S32 3x3", fwd, s32k3), ("wgrad S32 3x3", wgrad, s32k3)]
if TIME:
    s16 = make(16, 128, 1)
    CASES += [("wgrad+bias S64 1x1", wgrad_bias, s64), ("wgrad_bn-nobias S64", wgrad_bn_nob, s64),
              ("wgrad+bias S32 1x1", wgrad_bias, s32), ("wgrad_bn-nobias S32", wgrad_bn_nob, s32),
              ("wgrad S16 1x1", wgrad, s16), ("wgrad_bn S16 1x1", wgrad_bn, s16), ("wgrad_bn-nobias S16", wgrad_bn_nob, s16),
              ("fwd_bn S16 1x1", fwd_bn, s16), ("dgrad_bn S16 1x1", dgrad_bn, s16)]
torch.cuda.synchronize()
for name, fn, d in CASES:
    if TIME:
        for i in range(3):
            fn(d, i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for i in range(n):
            fn(d, i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        P, D, k = d["P"], d["D"], d["k"]
        byt = 4 * 2 * P * D + (4 * P * D if ("fwd_bn" in name or "dgrad" in name) else 0)
        print(f"{name:20s} {us:8.1f} us   {byt / us / 1e3:7.0f} GB/s algorithmic   {2.0 * P * D * D * k * k / us / 1e6:6.1f} TF/s")
    else:
        for i in range(NREP):
            fn(d, i)
        torch.cuda.synchronize()
print("done")
