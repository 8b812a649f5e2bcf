"""Micro-benchmark of the conv kernels through the C-ABI (GPU box only): time per launch with CUDA
events, rotating over several operand sets so that the working set exceeds L2 where it would in the
real step.  Usage: python tools/bench_conv.py [math]"""
import ctypes as C, importlib, os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
pkg = importlib.import_module("dl-normalizing-flows_b200")
lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
DEV = "cuda"
B = int(os.environ.get("B", "256"))
math = 1 if (len(sys.argv) < 2 or sys.argv[1] == "tf32") else 0
pad = lambda v, m: (v + m - 1) // m * m
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=None):
    n = n or NITER
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print(f"{'shape':26s} {'kernel':10s} {'us':>8s} {'GB/s':>7s} {'TF/s':>6s} {'ideal_us':>8s}")
SHAPES = [(64, 32, 1), (64, 32, 3), (32, 64, 1), (32, 64, 3), (16, 128, 1), (16, 128, 3),
          (8, 256, 1), (8, 256, 3), (4, 512, 1), (4, 512, 3)]
if os.environ.get("SHAPES"):
    SHAPES = [tuple(int(v) for v in t.split(",")) for t in os.environ["SHAPES"].split(";")]
NITER = int(os.environ.get("NITER", "20"))
for S, D, k in SHAPES:
    P = B * S * S
    kpad, npad = pad(D, 32), pad(D, 16)
    nset = max(2, min(6, int(300e6 // (P * D * 4)) + 1))
    xs = [torch.randn(B, S, S, kpad, device=DEV) for _ in range(nset)]
    ys = [torch.empty(B, S, S, kpad, device=DEV) for _ in range(nset)]
    rs = [torch.randn(B, S, S, kpad, device=DEV) for _ in range(nset)]
    w = torch.randn(k * k, npad, kpad, device=DEV) * 0.05
    dw = torch.zeros(k * k, npad, kpad, device=DEV)
    bias = torch.randn(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    stats = torch.zeros(2 * D, dtype=torch.float64, device=DEV)
    byt = 4 * (2 * P * D + k * k * npad * kpad)
    fl = 2.0 * P * D * D * k * k
    ideal = max(byt / 6547e3, fl / 685e6)
    cases = {
        "fwd": lambda i: check(lib.rnvp_conv_forward(ptr(xs[i % nset]), ptr(w), None, None, ptr(ys[i % nset]), None, B, S, kpad, D, npad, k, kpad, math, st)),
        "fwd+st": lambda i: check(lib.rnvp_conv_forward(ptr(xs[i % nset]), ptr(w), None, None, ptr(ys[i % nset]), ptr(stats), B, S, kpad, D, npad, k, kpad, math, st)),
        "fwd+b+r+s": lambda i: check(lib.rnvp_conv_forward(ptr(xs[i % nset]), ptr(w), ptr(bias), ptr(rs[i % nset]), ptr(ys[i % nset]), ptr(stats), B, S, kpad, D, npad, k, kpad, math, st)),
        "wgrad": lambda i: check(lib.rnvp_conv_wgrad(ptr(xs[i % nset]), ptr(rs[i % nset]), ptr(dw), ptr(db), B, S, kpad, D, npad, k, kpad, math, st)),
    }
    cases["wgrad-nob"] = lambda i: check(lib.rnvp_conv_wgrad(ptr(xs[i % nset]), ptr(rs[i % nset]), ptr(dw), None, B, S, kpad, D, npad, k, kpad, math, st))
    for name, fn in cases.items():
        us = timeit(fn)
        extra = 4 * P * D if "r" in name.split("+") else 0
        print(f"S{S} {D}->{D} {k}x{k}".ljust(26) + f" {name:10s} {us:8.1f} {(byt + extra) / us / 1e3:7.0f} {fl / us / 1e6:6.1f} {ideal:8.1f}")
    del xs, ys, rs
    torch.cuda.empty_cache()
