#!/usr/bin/env python
"""Benchmark of the RealNVP hot path (BASELINE.json: train imgs/s, RealNVP 64x64x3, base 32, 4 blocks).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun by the driver)
  python bench.py --impl reference ...                   (the reference algorithm on the host cores)

One "step" = one training step on a batch of synthetic dequantised images: logit transform, forward
log-likelihood through the multi-scale coupling stack, loss, backward, Adam update.

  value : device-timed throughput with the uint8 batch already resident in HBM
  e2e   : the same step through the public API from pinned HOST buffers (uint8 H2D copy and the
          `.item()` read of the loss inside the timed region), i.e. what train.py:176-200 does
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

CFG = dict(channels=3, image=64, base_dim=32, res_blocks=4, num_scales=5)
GFLOP_FWD_PER_IMG = 11.9634          # s/t convs only, multiply-add = 2
METRIC = "train imgs/s RealNVP 64x64x3 (fwd log-lik + bwd + Adam)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE config 2: 256)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: the WHOLE job's batch, split evenly over the GPUs (BASELINE configs[3]: 4096 "
                         "samples, configs[4]: global batch 2048); overrides --batch")
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32", "tf32x3"],
                    help="tf32: tcgen05 TF32 (default, the benchmarked tier); tf32x3: 3xTF32 split operands on the tensor cores "
                         "(fp32-accurate); fp32: CUDA-core fp32")
    ap.add_argument("--mode", default="train", choices=["train", "sample"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prof", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true",
                    help="skip timing the unmodified reference's own PyTorch-eager CUDA path on this GPU "
                         "(gpu_eager_baseline; 1-GPU runs only, needs oracle/_ref)")
    ap.add_argument("--config", default="A", choices=["A", "c3"],
                    help="A: RealNVP 64x64x3, base 32, 4 blocks, 5 scales (BASELINE configs[1], the judged line); "
                         "c3: the 32x32 two-scale variant, base 64, 8 blocks (BASELINE configs[2], batch 512)")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: rnvp_optim.Adam (one launch); torch: torch.optim.Adam(fused=True)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# the reference on the host cores: its own code from oracle/_ref (kind "reference"), else the oracle port
# ------------------------------------------------------------------------------------------
def _ref_modules(cpu):
    """(flow_realnvp, utils) of the UNMODIFIED reference (byte-compiled into oracle/_ref by oracle/build_ref.py),
    or None when it was not built (then the oracle port is timed instead)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_loader
        if ref_loader.available():
            f, _m, u = ref_loader.load(cpu=cpu)
            return f, u
    except Exception as e:                      # pragma: no cover - diagnostic only
        print(f"[bench] reference modules unavailable: {e}", file=sys.stderr)
    return None


def ref_train_step_factory(batch, device="cpu"):
    """train.py:176-200 replayed literally on the reference's own RealNVP: zero_grad, logit_transform on the CPU
    batch (train.py:187), .to(device), model(x), loss, .item(), backward, Adam step."""
    import torch
    cpu = device == "cpu"
    mods = _ref_modules(cpu)
    if cpu:
        torch.set_num_threads(os.cpu_count())
    if mods is None:
        if not cpu:
            return None, None, None
        return cpu_port_train_step_factory(batch) + ("port",)
    f, u = mods
    dev = torch.device(device)
    torch.manual_seed(0)
    prior = torch.distributions.Normal(torch.tensor(0., device=dev), torch.tensor(1., device=dev), validate_args=False)
    hps = u.Hyperparameters(CFG["base_dim"], CFG["res_blocks"], True, True, True, True)
    assert CFG["num_scales"] == 5, "the in-tree reference hard-codes five scales"
    model = f.RealNVP(CFG["channels"], CFG["image"], prior, hps).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-5)            # train.py:134
    g = torch.Generator().manual_seed(0)
    x_img = torch.randint(0, 256, (batch, CFG["channels"], CFG["image"], CFG["image"]), generator=g,
                          dtype=torch.uint8).float() / 255.0

    def step():
        opt.zero_grad()
        x, logdet = u.logit_transform(x_img)
        x, logdet = x.to(dev), logdet.to(dev)
        logll, weight_scale = model(x)
        logll = (logll + logdet).mean()
        loss = -logll + 5e-5 * weight_scale
        v = logll.item()
        loss.backward()
        opt.step()
        return v
    return step, (torch.get_num_threads() if cpu else 0), "reference"


def ref_sample_step_factory(batch, device="cpu"):
    import torch
    cpu = device == "cpu"
    mods = _ref_modules(cpu)
    if cpu:
        torch.set_num_threads(os.cpu_count())
    if mods is None:
        if not cpu:
            return None, None, None
        return cpu_port_sample_step_factory(batch) + ("port",)
    f, u = mods
    dev = torch.device(device)
    torch.manual_seed(0)
    prior = torch.distributions.Normal(torch.tensor(0., device=dev), torch.tensor(1., device=dev), validate_args=False)
    hps = u.Hyperparameters(CFG["base_dim"], CFG["res_blocks"], True, True, True, True)
    model = f.RealNVP(CFG["channels"], CFG["image"], prior, hps).to(dev)
    model.eval()

    def step():
        with torch.no_grad():                                                        # train.py:253-257
            imgs, _ = u.logit_transform(model.sample(size=batch), reverse=True)
            return float(imgs.mean())
    return step, (torch.get_num_threads() if cpu else 0), "reference"


def cpu_port_train_step_factory(batch):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import realnvp_oracle as O
    torch.set_num_threads(os.cpu_count())
    st = O.random_state(CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"], seed=0)
    for k, v in st.items():
        if O.is_trainable(k) and v.is_floating_point():
            v.requires_grad_(True)
    ora = O.RealNVPOracle(st, CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"])
    params = [v for k, v in st.items() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4, weight_decay=5e-5)
    x_img = O.synthetic_images(batch, CFG["channels"], CFG["image"], seed=0)

    def step():
        opt.zero_grad()
        noise = torch.rand(x_img.shape)
        x, ld = O.logit_forward(x_img, noise)
        ll, ws = ora.forward(x)
        loss = -(ll + ld).mean() + 5e-5 * ws
        loss.backward()
        opt.step()
        return float(loss)
    return step, torch.get_num_threads()


def cpu_port_sample_step_factory(batch):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import realnvp_oracle as O
    torch.set_num_threads(os.cpu_count())
    st = O.random_state(CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"], seed=0)
    ora = O.RealNVPOracle(st, CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"])
    ora.training = False

    def step():
        with torch.no_grad():
            z = torch.randn(batch, CFG["channels"], CFG["image"], CFG["image"])
            return float(O.logit_inverse(ora.g(z)).mean())
    return step, torch.get_num_threads()


def time_cpu(step, warmup, steps, budget_s=25.0):
    for _ in range(warmup):
        step()
    ts = []
    t_all = time.time()
    for _ in range(steps):
        t0 = time.time()
        step()
        ts.append(time.time() - t0)
        if time.time() - t_all > budget_s:
            break
    ts.sort()
    return ts[len(ts) // 2], len(ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 64                                           # BASELINE configs[0]: batch 64 on the CPU
    factory = ref_train_step_factory if args.mode == "train" else ref_sample_step_factory
    if args.config != "A":                               # the two-scale variant is not in the reference tree
        factory = (lambda b: cpu_port_train_step_factory(b) + ("port",)) if args.mode == "train" else \
                  (lambda b: cpu_port_sample_step_factory(b) + ("port",))
    step, threads, kind = factory(batch)
    warm = min(args.warmup, 1)
    med, n = time_cpu(step, warm, max(1, min(args.steps, 5)), budget_s=120.0)
    v = batch / med
    unit = "img/s"
    what = ("the UNMODIFIED reference (byte-compiled into oracle/_ref), train.py:176-200 replayed on its RealNVP"
            if kind == "reference" else "oracle port of the reference algorithm, torch CPU ops")
    out = {"metric": METRIC if args.mode == "train" else "sample imgs/s RealNVP 64x64x3", "value": v, "unit": unit,
           "n_gpus": args.gpus, "steps": n, "warmup": warm, "ms_per_step": med * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": "RealNVP 64x64x3->4x4x48, 4 res-blocks, base-dim 32 (BASELINE configs[0])",
                      "batch_per_step": batch, "note": what},
           "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": kind,
                            "sample": f"{n} steps of batch {batch} ({args.mode})"},
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _RECORD.append(json.dumps(out))


# ------------------------------------------------------------------------------------------
# the B200 path
# ------------------------------------------------------------------------------------------
def conv_alg_bytes_and_flops(kind, S, taps, cin, cout, B, fused_reduce=True):
    """Algorithmic traffic / flops of one profiled launch class (fp32 NHWC, every operand once)."""
    pad = lambda v, m: (v + m - 1) // m * m
    P = B * S * S
    if kind == 0:               # conv: read input rows, write output rows, read weights (a residual read, where the
        byt = 4 * (P * pad(cin, 32) + P * cout + taps * pad(cout, 16) * pad(cin, 32))     # layer has one, is not counted)
        flops = 2.0 * P * cin * cout * taps
    elif kind == 1:             # dgrad: as conv, plus the second input every dgrad of the stack reads -- the saved pre-BN
        # activation of the fused ReLU/BN-backward epilogue, or the running gradient it accumulates into
        byt = 4 * (P * pad(cin, 32) + 2 * P * cout + taps * pad(cout, 16) * pad(cin, 32))
        flops = 2.0 * P * cin * cout * taps
    elif kind == 2:             # wgrad: read x rows and dy rows, write dw
        byt = 4 * (P * pad(cin, 32) + P * pad(cout, 32) + taps * pad(cout, 16) * pad(cin, 32))
        flops = 2.0 * P * cin * cout * taps
    elif kind == 3:             # bn+relu apply: read x, write h
        byt, flops = 4 * 2 * P * cin, 0.0
    elif fused_reduce:          # bn backward, tensor-core tier: the reduce rides in the dgrad epilogue; only the apply
        byt, flops = 4 * 3 * P * cin, 0.0          # kernel runs (read g, x; write dx)
    else:                       # bn backward = reduce (read g, x; write g*mask) + apply (read g, x; write dx)
        byt, flops = 4 * 6 * P * cin, 0.0
    return byt, flops


def kernel_name_of(kind, taps, cin, cout, math, xform):
    """The CUDA kernel (as ncu names it) a profiled launch class runs in."""
    if kind in (0, 1):
        if math == "fp32":
            return "conv_fwd_fp32_kernel"
        bn = 32 if cout <= 32 else (64 if cout <= 64 else 128)
        x3 = 1 if math == "tf32x3" else 0            # <BN, transform warps, coupling epilogue, 3xTF32>
        return f"conv_fwd_tf32_kernel<{bn},{1 if (x3 or (kind == 0 and xform)) else 0},0,{x3}>"
    if kind == 2:
        return "conv_wgrad_tf32_kernel" if math != "fp32" else "conv_wgrad_fp32_kernel"
    return "bn_relu_kernel" if kind == 3 else "bn_bwd_apply_kernel"


def gpu_eager_baseline(args, dev, B):
    """The UNMODIFIED reference on cuda:0 through its own eager path (cuDNN / ATen), train.py:176-200 replayed at the
    bench's batch: 3 warm-up + 5 timed steps per setting.  torch's default lets cuDNN use TF32 for convolutions
    (torch.backends.cudnn.allow_tf32 = True), which is what the reference gets on this GPU; the strict-fp32 setting
    is timed beside it."""
    import gc
    import torch
    out = {"batch": B, "steps": 5, "warmup": 3, "unit": "img/s", "kind": "reference",
           "what": "reference RealNVP (oracle/_ref), PyTorch eager on the same GPU, train.py:176-200 incl. CPU logit_transform, "
                   ".item() and torch.optim.Adam" if args.mode == "train" else "reference model.sample + inverse logit, eager"}
    for tag, allow in (("cudnn_tf32_default", True), ("fp32_strict", False)):
        torch.backends.cudnn.allow_tf32 = allow
        try:
            factory = ref_train_step_factory if args.mode == "train" else ref_sample_step_factory
            step, _t, kind = factory(B, device=str(dev))
            if step is None:
                return {"unavailable": "oracle/_ref not built"}
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                step()
            torch.cuda.synchronize()
            out[tag] = B * 5 / (time.perf_counter() - t0)
        except Exception as e:                                    # e.g. out of memory at a large batch
            out[tag] = None
            out[tag + "_error"] = str(e)[:200]
        finally:
            step = None
            gc.collect()
            torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    return out


def dp_consistency_check(pkg, model, net, dev, rank, world):
    """Data-parallel result == single process on the concatenated batch (per-sample log-likelihood of this rank's
    shard and the averaged gradient), at a small batch so that rank 0 can run the global batch alone."""
    import copy
    import torch
    import torch.distributed as dist
    b = 4
    g = torch.Generator().manual_seed(77)
    xs = torch.randn(world * b, CFG["channels"], CFG["image"], CFG["image"], generator=g).to(dev)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    # The check runs in the fp32 tier: at a global batch of 8 the TF32 gradient of this 28-coupling train-mode stack is
    # chaotic (two valid summation orders differ by O(0.1)), which would make the comparison a coin toss;
    # the data-parallel plumbing under test (shards, statistic exchange, gradient all-reduce) is the same in every tier.
    math_was = "tf32" if net.engine().math == pkg.rnvp_cabi.MATH_TF32 else ("fp32" if net.engine().math == pkg.rnvp_cabi.MATH_FP32 else "tf32x3")
    net.set_math("fp32")
    net.train()
    net.zero_grad(set_to_none=True)
    ll, ws = model(xs[rank * b:(rank + 1) * b].contiguous())
    (-ll.mean() + 5e-5 * ws).backward()
    key = next(k for k, p in net.named_parameters() if k.endswith("out_block.2.conv.weight_v"))
    gdp = dict(net.named_parameters())[key].grad.detach().clone()
    gathered = [torch.zeros_like(ll) for _ in range(world)]
    dist.all_gather(gathered, ll.detach())
    ok, detail = True, {}
    if rank == 0:
        prior = torch.distributions.Normal(torch.tensor(0., device=dev), torch.tensor(1., device=dev), validate_args=False)
        ref = pkg.RealNVP(CFG["channels"], CFG["image"], prior,
                          pkg.Hyperparameters(CFG["base_dim"], CFG["res_blocks"], True, True, True, True),
                          **({} if CFG["num_scales"] == 5 else {"num_scales": CFG["num_scales"]})).to(dev)
        ref.load_state_dict(sd)
        ref.set_math("fp32")
        ref.train()
        ll1, ws1 = ref(xs)
        (-ll1.mean() + 5e-5 * ws1).backward()
        g1 = dict(ref.named_parameters())[key].grad
        e_ll = float((torch.cat(gathered) - ll1.detach()).abs().max() / ll1.detach().abs().max())
        e_g = float((gdp - g1).norm() / g1.norm())
        detail = {"ll_rel": e_ll, "grad_rel_l2": e_g, "global_batch": world * b}
        detail["math"] = "fp32"
        ok = e_ll < 1e-4 and e_g < 2e-2
        del ref
    net.set_math(math_was)
    net.load_state_dict(sd)                        # running statistics back to where they were
    net.zero_grad(set_to_none=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    return {"status": "ok" if int(flag) else "MISMATCH", **detail}


def run_b200(args):
    # NCCL_DEBUG is left as the environment has it: stdout is redirected to stderr for the duration of the run
    # (see main), so NCCL's INFO lines cannot break the one-JSON-line contract and the driver can count ranks
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("dl-normalizing-flows_b200")
    cabi = pkg.rnvp_cabi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cabi.check(cabi.lib.rnvp_device_ok())
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} GPUs")
        args.batch = args.global_batch // world
    B = args.batch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

    torch.manual_seed(0)
    prior = torch.distributions.Normal(torch.tensor(0., device=dev), torch.tensor(1., device=dev), validate_args=False)
    model = pkg.RealNVP(CFG["channels"], CFG["image"], prior,
                        pkg.Hyperparameters(CFG["base_dim"], CFG["res_blocks"], True, True, True, True),
                        **({} if CFG["num_scales"] == 5 else {"num_scales": CFG["num_scales"]})).to(dev)
    model.set_math(args.math)
    if world > 1 and args.mode == "train":              # sampling shards by batch with no communication
        import rnvp_dp
        model = rnvp_dp.DataParallel(model)
    net = model.module if (world > 1 and args.mode == "train") else model
    if args.optimizer == "fused":
        opt = pkg.rnvp_optim.Adam(model, lr=5e-4, weight_decay=5e-5)           # train.py:134 hyper-parameters
    else:
        opt = torch.optim.Adam(net.parameters(), lr=5e-4, weight_decay=5e-5, fused=True)

    g = torch.Generator().manual_seed(1234 + rank)
    host_u8 = torch.randint(0, 256, (B, CFG["channels"], CFG["image"], CFG["image"]), generator=g,
                            dtype=torch.uint8).pin_memory()
    dev_u8 = host_u8.to(dev)

    def train_step(x_u8):
        opt.zero_grad(set_to_none=args.optimizer == "torch")
        x, logdet = pkg.logit_transform(x_u8)
        ll, wscale = model(x)
        loss = -(ll + logdet).mean() + 5e-5 * wscale
        loss.backward()
        opt.step()
        return loss

    sample_out = {}

    def sample_step(_):
        with torch.no_grad():
            imgs, _ = pkg.logit_transform(net.sample(B), reverse=True)
        sample_out["imgs"] = imgs
        return imgs.mean()

    net.train(args.mode == "train")
    if args.mode == "sample":
        torch.manual_seed(4321 + rank)                      # every rank draws its own z
        # converged running statistics first (a sampler is used after training); the training workspace of the
        # full sampling batch would not fit, so these steps use at most 256 images
        net.train()
        for _ in range(3):
            train_step(dev_u8[:256])
        net.eval()
        opt.zero_grad(set_to_none=True)
        net.engine().release_workspaces()
    step = train_step if args.mode == "train" else sample_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def note(msg):
        if os.environ.get("RNVP_BENCH_VERBOSE"):
            print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

    # ---- warm-up -------------------------------------------------------------------------------
    note("warm-up")
    for i in range(max(args.warmup, 3)):
        step(dev_u8)
        note(f"warm-up step {i} enqueued")
    barrier()
    note("warm-up done")
    dp_check = None
    if world > 1 and args.mode == "train":
        dp_check = dp_consistency_check(pkg, model, net, dev, rank, world)
        opt.zero_grad(set_to_none=args.optimizer == "torch")
        barrier()

    # ---- device-resident timing -------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = cabi.lib.rnvp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dev_u8)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = cabi.lib.rnvp_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end from pinned host memory ----------------------------------------------------------
    host_imgs = (torch.empty(B, CFG["channels"], CFG["image"], CFG["image"], dtype=torch.float32).pin_memory()
                 if args.mode == "sample" else None)

    def e2e_step():
        if args.mode == "train":
            return float(step(host_u8.to(dev, non_blocking=True)))    # the caller's logll.item() (train.py:196)
        step(None)
        host_imgs.copy_(sample_out["imgs"], non_blocking=True)        # a sampler's result is the images (train.py:257-259)
        torch.cuda.synchronize()
        return float(host_imgs[0, 0, 0, 0])
    for _ in range(2):                                        # this path's own warm-up (allocator, pinned-copy queue)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te)
    h2d = host_u8.numel() if args.mode == "train" else 0
    d2h = 4 if args.mode == "train" else host_imgs.numel() * 4

    # ---- per-kernel-class event timing: the roofline line ---------------------------------------------
    roof = None
    classes = []
    by_kind = {}
    nprof = 2
    if not args.no_prof:
        # every rank runs the profiled steps (under data parallelism a step is a collective)
        cabi.lib.rnvp_prof_enable(1)
        for _ in range(nprof):
            step(dev_u8)
        rows = (C.c_double * (7 * 512))()
        n = cabi.lib.rnvp_prof_collect(rows, 512)
        cabi.lib.rnvp_prof_enable(0)
        barrier()
    if rank == 0 and not args.no_prof:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                mp = json.load(f)
            peaks = {"hbm_gbs": mp["hbm_gbs"], "bf16_tflops": mp["bf16_tflops_sustained"], "src": "measured"}
        except Exception:
            pass
        names = {0: "conv", 1: "dgrad", 2: "wgrad", 3: "bn_relu", 4: "bn_bwd"}
        tot = 0.0
        for i in range(max(n, 0)):
            kind, S, taps, cin, cout, cnt, tms = [rows[7 * i + j] for j in range(7)]
            tot += tms
            classes.append(dict(kind=int(kind), S=int(S), taps=int(taps), cin=int(cin), cout=int(cout),
                                launches=int(cnt), ms=tms))
        classes.sort(key=lambda r: -r["ms"])
        if os.environ.get("RNVP_BENCH_CLASSES"):          # full per-class table (the JSON line keeps the top 16)
            with open(os.environ["RNVP_BENCH_CLASSES"], "w") as f:
                json.dump({"batch": B, "profiled_steps": nprof, "classes": classes}, f, indent=0)
        for r in classes:
            by_kind[names[r["kind"]]] = by_kind.get(names[r["kind"]], 0.0) + r["ms"] / nprof
        xform_on = args.math != "fp32" and os.environ.get("RNVP_XFORM", "1") != "0"
        groups = {}
        for r in classes:
            byt, fl = conv_alg_bytes_and_flops(r["kind"], r["S"], r["taps"], r["cin"], r["cout"], B,
                                               fused_reduce=args.math == "tf32")
            r["alg_bytes"], r["alg_flops"] = byt, fl
            r["kernel"] = kernel_name_of(r["kind"], r["taps"], r["cin"], r["cout"], args.math, xform_on)
            gsum = groups.setdefault(r["kernel"], dict(ms=0.0, launches=0, bytes=0.0, flops=0.0, roof_ms=0.0))
            gsum["ms"] += r["ms"]; gsum["launches"] += r["launches"]
            gsum["bytes"] += byt * r["launches"]; gsum["flops"] += fl * r["launches"]
            # time this class would take at ITS OWN binding roof (a kernel name covers HBM-bound and tensor-bound shapes)
            gsum["roof_ms"] += r["launches"] * 1e3 * max(byt / (peaks["hbm_gbs"] * 1e9),
                                                         fl / (peaks["bf16_tflops"] * 0.5 * 1e12))
        if groups:
            # the dominant kernel = the CUDA kernel (ncu name) with the largest share of the profiled kernel time;
            # achieved = its algorithmic bytes (flops) per launch / its average launch duration, over all its launches
            kname, gk = max(groups.items(), key=lambda kv: kv[1]["ms"])
            dur = gk["ms"] / gk["launches"] / 1e3
            byt, fl = gk["bytes"] / gk["launches"], gk["flops"] / gk["launches"]
            gbs, tfs = byt / dur / 1e9, fl / dur / 1e12
            tf32_peak = peaks["bf16_tflops"] * 0.5            # kind::tf32 runs at half the bf16 rate
            t_hbm, t_tc = byt / (peaks["hbm_gbs"] * 1e9), fl / (tf32_peak * 1e12)
            if t_hbm >= t_tc:
                roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": None}
            else:
                roof = {"bound": "tensor", "achieved": tfs, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": tfs / tf32_peak, "traffic": None,
                        "peak_note": "tf32 tensor peak taken as half the measured sustained bf16 cuBLAS rate"}
            roof["kernel"] = kname
            roof["launches_per_step"] = gk["launches"] // nprof
            roof["avg_us"] = dur * 1e6
            roof["share_of_profiled_kernel_time"] = gk["ms"] / tot
            roof["peak_source"] = peaks["src"]
            roof["schedule"] = ("durations taken with every kernel on one stream (rnvp_prof_enable switches the "
                                "wgrad side stream off), CUDA events around each launch")
            roof["algorithmic_bytes_per_launch"] = byt
            roof["algorithmic_flops_per_launch"] = fl
            roof["hbm_frac_of_this_kernel"] = gbs / peaks["hbm_gbs"]
            roof["tensor_frac_of_this_kernel"] = tfs / tf32_peak
            # the same kernel with every launch class held against its own binding roof (HBM for the S >= 16 shapes,
            # the TF32 tensor roof for the 3x3 / deep ones): sum of roofline times / sum of measured times
            roof["frac_vs_own_roof_per_class"] = gk["roof_ms"] / gk["ms"]
            try:                                          # DRAM bytes per launch from the committed ncu capture
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    roof["traffic"] = json.load(f).get(kname)
            except Exception:
                pass
            roof["by_kernel"] = {k: {"share": v["ms"] / tot, "launches_per_step": v["launches"] // nprof,
                                     "avg_us": v["ms"] / v["launches"] * 1e3,
                                     "hbm_frac": v["bytes"] / (v["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"],
                                     "tensor_frac": v["flops"] / (v["ms"] / 1e3) / 1e12 / tf32_peak,
                                     "frac_vs_own_roof_per_class": v["roof_ms"] / v["ms"]}
                                 for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
            # whole-step roofline on SURVEY.md 8(d) algorithmic bytes: conv activations in + out once each
            # (41.436 M elements per image and pass, fp32 storage), training = 2.5 passes
            if args.config == "A":
                step_bytes = 41.436e6 * 4 * (2.5 if args.mode == "train" else 1.0) * B
                roof["step"] = {"algorithmic_bytes_per_step": step_bytes,
                                "ms_at_hbm_roof": step_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3,
                                "frac": step_bytes / (peaks["hbm_gbs"] * 1e9) / (ms / args.steps / 1e3),
                                "tensor_frac": value / world * (3 if args.mode == "train" else 1) * GFLOP_FWD_PER_IMG / 1e3 / tf32_peak}

    # ---- the reference's own PyTorch-eager CUDA path on this GPU ("the bar on the box") ---------
    eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager and args.config == "A":
        net.engine().release_workspaces()                      # hand the training workspace back before the eager model
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args, dev, B)

    # ---- CPU baseline (bounded sample of the same workload) ---------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        if args.config == "A":
            factory = ref_train_step_factory if args.mode == "train" else ref_sample_step_factory
            cstep, threads, kind = factory(64)
        else:
            factory = cpu_port_train_step_factory if args.mode == "train" else cpu_port_sample_step_factory
            (cstep, threads), kind = factory(64), "port"
        med, n = time_cpu(cstep, 1, 3, budget_s=25.0)
        cpu = {"value": 64 / med, "unit": "img/s", "cores": threads, "kind": kind,
               "sample": f"{n} {args.mode} steps of batch 64 (BASELINE configs[0]) on the host cores"}

    if rank == 0:
        tr_flops = 3 * GFLOP_FWD_PER_IMG if args.mode == "train" else GFLOP_FWD_PER_IMG
        out = {"metric": METRIC if args.mode == "train" else "sample imgs/s RealNVP 64x64x3",
               "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
               "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
               "config": {"workload": ("RealNVP 32x32x3 -> 16x16x6, 8 res-blocks / 64 features (BASELINE configs[2])"
                                       if args.config == "c3" else
                                       f"RealNVP 64x64x3, 4 res-blocks / 32 features, batch {B} per GPU "
                                       "(BASELINE configs[1])" if args.mode == "train" else
                                       "RealNVP 64x64x3 inverse sampling (BASELINE configs[3])"),
                          "batch_per_gpu": B, "global_batch": B * world, "mode": args.mode,
                          "parallelism": f"dp{world}" if world > 1 else "single",
                          **({"bn_stat_exchange": model.stat_exchange, "grad_allreduce": "nccl, bucketed, overlapped"}
                             if (world > 1 and args.mode == "train") else {}),
                          "l2": "inputs larger than L2: the per-step working set (tens of GB of activations) exceeds the 126 MB L2",
                          "optimizer": ("rnvp_optim.Adam (one fused launch, clears the gradients)"
                                        if args.optimizer == "fused" else "torch.optim.Adam(fused=True)") +
                                       " inside the step"},
               "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(launches),
               "achieved_tflops_algorithmic": value * tr_flops / 1e3,
               "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
               **({"dp_check": dp_check} if world > 1 else {}),
               "kernel_time_ms_by_kind_per_step": by_kind, "kernel_classes": classes[:16]}
        _RECORD.append(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON record: anything a library prints there (NCCL's version banner,
    # torchrun notices) is sent to stderr instead by swapping the descriptors for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(args)
    finally:
        sys.stdout.flush()
        # the record goes straight to the real stdout; fd 1 STAYS pointed at stderr, so whatever a library still
        # prints at teardown (NCCL_DEBUG=INFO's communicator-destroy lines) cannot follow the JSON line
        if _RECORD:
            os.write(real_stdout, (_RECORD[-1] + "\n").encode())
        os.close(real_stdout)


_RECORD = []


def _main(args):
    global GFLOP_FWD_PER_IMG, METRIC
    if args.config == "c3":                          # SURVEY.md 8d "Config 3": 10 couplings, 8.2215 GFLOP/img forward
        CFG.update(image=32, base_dim=64, res_blocks=8, num_scales=2)
        GFLOP_FWD_PER_IMG = 8.2215
        METRIC = "train imgs/s RealNVP 32x32x3 two-scale (fwd log-lik + bwd + Adam)"
    wd = int(os.environ.get("RNVP_BENCH_WATCHDOG", "0"))
    if wd:
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
