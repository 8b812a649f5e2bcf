"""Wall time of ONE coupling (training forward + backward, eval forward, inverse) per scale of cfg A, through the
stand-alone coupling modules -- the real schedule (PDL, side streams), not the per-kernel event profile.
GPU box only.  Usage: [B=256] [MATH=tf32] python tools/bench_coupling.py
Prints one line per (kind, S, D): ms per call and the number of kernels launched."""
import importlib
import os
import sys
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch

pkg = importlib.import_module("dl-normalizing-flows_b200")
lib = pkg.rnvp_cabi.lib
B = int(os.environ.get("B", "256"))
MATH = os.environ.get("MATH", "tf32")
R = int(os.environ.get("R", "4"))
NIT = int(os.environ.get("NITER", "10"))
# (kind, C, S, D, couplings of this shape in cfg A)
SHAPES = [("ckbd", 3, 64, 32, 3), ("chan", 12, 32, 64, 3), ("ckbd", 6, 32, 64, 3), ("chan", 24, 16, 128, 3),
          ("ckbd", 12, 16, 128, 3), ("chan", 48, 8, 256, 3), ("ckbd", 24, 8, 256, 3), ("chan", 96, 4, 512, 3),
          ("ckbd", 48, 4, 512, 4)]
if os.environ.get("ONLY"):
    keep = [int(v) for v in os.environ["ONLY"].split(",")]
    SHAPES = [s for i, s in enumerate(SHAPES) if i in keep]
pkg.set_default_math(MATH)


def timed(fn, n=NIT):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    l0 = lib.rnvp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (lib.rnvp_launch_count() - l0) // n


tot = {"train": 0.0, "eval": 0.0, "inv": 0.0}
print(f"B={B} math={MATH}  {'shape':22s} {'train f+b ms':>12s} {'launches':>8s} {'eval ms':>8s} {'inv ms':>8s}   x couplings")
for kind, C, S, D, cnt in SHAPES:
    hps = pkg.Hyperparameters(32, R, True, True, True, True)
    torch.manual_seed(0)
    mod = (pkg.CheckerboardAffineCoupling(C, D, S, 1.0, hps) if kind == "ckbd"
           else pkg.ChannelwiseAffineCoupling(C, D, 0.0, hps)).to("cuda")
    x = torch.randn(B, C, S, S, device="cuda")
    gy = torch.randn_like(x)

    def train_step():
        mod.train()
        xx = x.detach().requires_grad_(True)
        y, J = mod(xx)
        (y * gy).sum().add(J.sum()).backward()

    def eval_step():
        mod.eval()
        with torch.no_grad():
            mod(x)

    def inv_step():
        mod.eval()
        with torch.no_grad():
            mod(x, reverse=True)

    t_tr, n_tr = timed(train_step)
    t_ev, _ = timed(eval_step)
    t_in, _ = timed(inv_step)
    tot["train"] += cnt * t_tr
    tot["eval"] += cnt * t_ev
    tot["inv"] += cnt * t_in
    print(f"{'':14s}{kind} C{C} S{S} D{D}".ljust(37) + f"{t_tr:12.3f} {n_tr:8d} {t_ev:8.3f} {t_in:8.3f}   x{cnt}")
    del mod, x, gy
    torch.cuda.empty_cache()
print(f"sum over the 28 couplings of cfg A: train {tot['train']:.2f} ms, eval {tot['eval']:.2f} ms, inverse {tot['inv']:.2f} ms "
      f"(stand-alone calls: each includes the NCHW<->NHWC transposes, the weight norm of its 19 convs and the torch-side autograd glue)")
