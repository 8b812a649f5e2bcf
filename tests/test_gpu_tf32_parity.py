"""-m gpu: the tensor-core (TF32) tier, which every benchmark number is quoted on.

Three kinds of gates:
  * against the fp32 reference / oracle: per-sample log-likelihood and log-det <= 1e-3 (north star), train mode and
    eval mode with converged running statistics;
  * against the TF32-EMULATING oracle (oracle/realnvp_oracle.py, emulate_tf32=True: every conv-MMA operand rounded
    to nearest TF32 by its producer, fp32 accumulation, everything else fp32): this isolates kernel bugs from the
    cost of TF32 operands -- the only differences left are summation order and roundings flipped by 1e-7
    perturbations;
  * gradients: the end-to-end gradient of this 28-coupling stack amplifies perturbations by ~1e5 (the
    reference's own fp32-vs-fp64 floor is 5.5e-3), so under TF32 operands (2^-11) it is O(0.3 .. 0.6) away from the
    fp32 gradient for ANY implementation: the emulating oracle, an ideal round-to-nearest TF32, shows the same
    distance on the CPU (tools/diag_precision.py).  What is gated is therefore (a) that the CUDA tier is
    no further from fp32 than that ideal, (b) per-op and per-coupling VJPs, and (c) that training with the tier
    tracks training with the fp32 tier.
"""
import math
import os

import pytest
import torch

import realnvp_oracle as O
from _util import rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(pkg, c, state, math_mode):
    prior = torch.distributions.Normal(torch.tensor(0., device=DEV), torch.tensor(1., device=DEV), validate_args=False)
    hps = pkg.Hyperparameters(c["base_dim"], c["res_blocks"], True, True, True, True)
    kw = {} if c.get("num_scales", 5) == 5 else {"num_scales": c["num_scales"]}
    m = pkg.RealNVP(c["channels"], c["image"], prior, hps, **kw)
    m.load_state_dict(state, strict=True)
    m = m.to(DEV)
    m.set_math(math_mode)
    return m


def grad_distance(got, ref):
    num = den = dot = n1 = 0.0
    for k, b in ref.items():
        a, b = got[k].detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
        num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum())
        dot += float((a * b).sum()); n1 += float((a ** 2).sum())
    return (num / den) ** 0.5, dot / (n1 * den) ** 0.5


def oracle_fwd_bwd(st0, cfg, x, emu):
    ost = {k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st0.items()}
    ora = O.RealNVPOracle(ost, *cfg, emulate_tf32=emu)
    z, ld, lp = ora.log_prob_parts(x)
    ll = lp + ld
    (-(ll).mean() + 5e-5 * ora.weight_scale()).backward()
    return ll.detach(), ld.detach(), {k: v.grad for k, v in ost.items() if v.grad is not None}


CASES = {
    # name: (channels, image, base, R, L, B, scale, seed)
    "tiny_32px": (3, 32, 4, 2, 5, 4, 0.7, 3),
    "cfgA_b8_scale.7": (3, 64, 32, 4, 5, 8, 0.7, 0),
    "cfgA_b8_scale.2": (3, 64, 32, 4, 5, 8, 0.2, 0),
}


@pytest.mark.parametrize("name", list(CASES))
def test_tf32_tier_vs_fp32_and_emulated_oracle(pkg, name):
    ch, img, base, R, L, B, scale, seed = CASES[name]
    cfg = (ch, img, base, R, L)
    c = dict(channels=ch, image=img, base_dim=base, res_blocks=R, num_scales=L)
    st0 = O.random_state(ch, img, base, R, L, seed=seed, scale=scale)
    x_img = O.synthetic_images(B, ch, img, seed=seed)
    x, _ = O.logit_forward(x_img, torch.rand(x_img.shape, generator=torch.Generator().manual_seed(1)))
    ll_f, ld_f, g_f = oracle_fwd_bwd(st0, cfg, x, emu=False)
    ll_e, ld_e, g_e = oracle_fwd_bwd(st0, cfg, x, emu=True)
    m = build(pkg, c, st0, "tf32")
    m.train()
    ll, ws = m(x.to(DEV))
    (-(ll).mean() + 5e-5 * ws).backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    m2 = build(pkg, c, st0, "tf32")
    m2.train()
    _, ld, ll2 = m2.latent(x.to(DEV))
    # north-star gate vs the fp32 reference arithmetic
    assert rel(ll, ll_f) < 1e-3 and rel(ld, ld_f) < 1e-3, (rel(ll, ll_f), rel(ld, ld_f))
    # kernel-correctness gate vs the emulation of the tier's own arithmetic
    assert rel(ll, ll_e) < GATE_LL_EMU and rel(ld, ld_e) < GATE_LL_EMU, (rel(ll, ll_e), rel(ld, ld_e))
    d_cuda, cos_cuda = grad_distance(got, g_f)
    d_ideal, cos_ideal = grad_distance(g_e, g_f)
    d_emu, cos_emu = grad_distance(got, g_e)
    print(f"[{name}] ll vs fp32 {rel(ll, ll_f):.2e} vs emu {rel(ll, ll_e):.2e}; logdet vs fp32 {rel(ld, ld_f):.2e} vs emu "
          f"{rel(ld, ld_e):.2e}; grad: cuda-fp32 {d_cuda:.3f} (cos {cos_cuda:.4f}), ideal-tf32-fp32 {d_ideal:.3f} "
          f"(cos {cos_ideal:.4f}), cuda-emu {d_emu:.3f} (cos {cos_emu:.4f})")
    # the CUDA tier is no further from the fp32 gradient than an ideal round-to-nearest TF32 implementation
    assert d_cuda < 1.25 * d_ideal + 0.02, (d_cuda, d_ideal)
    # eval mode, converged running statistics, both oracles on the device model's buffers
    with torch.no_grad():
        for _ in range(30):
            m2(x.to(DEV))
    sd = {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()}
    m2.eval()
    with torch.no_grad():
        _, ld_d, ll_d = m2.latent(x.to(DEV))
    for emu, gate in ((False, 1e-3), (True, GATE_LL_EMU_EVAL)):
        oe = O.RealNVPOracle({k: v.clone() for k, v in sd.items()}, *cfg, emulate_tf32=emu)
        oe.training = False
        with torch.no_grad():
            _z, ld_o, lp_o = oe.log_prob_parts(x)
        print(f"[{name}] eval vs {'emu ' if emu else 'fp32'}: ll {rel(ll_d, lp_o + ld_o):.2e} logdet {rel(ld_d, ld_o):.2e}")
        assert rel(ll_d, lp_o + ld_o) < gate and rel(ld_d, ld_o) < gate, (emu, rel(ll_d, lp_o + ld_o), rel(ld_d, ld_o))


# Measured on B200 (profiles/r02_parity_diag.log): 3e-5 ... 2.5e-4.  Two implementations of the SAME rounding scheme
# still differ by summation order (1e-7), which flips individual TF32 roundings, and the stack amplifies those flips
# like any other perturbation -- the ideal emulation itself sits 2e-5 ... 1.6e-4 away from fp32 at these points.
GATE_LL_EMU = 5e-4
GATE_LL_EMU_EVAL = 5e-4


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_tf32_coupling_vs_emulated_oracle(pkg, golden_dir, mode):
    """Stand-alone couplings of both kinds (tests/golden/couplings.pt shapes and states): forward, inverse and VJP
    of the tensor-core tier against the TF32-emulating oracle on the same state."""
    fix = torch.load(os.path.join(golden_dir, "couplings.pt"))
    pkg.set_default_math("tf32")
    worst = {}
    for tag, case in fix.items():
        kind, C, S, D, cfg, R = case["kind"], case["C"], case["S"], case["D"], case["cfg"], case["R"]
        hps = pkg.Hyperparameters(8, R, True, True, True, True)
        mod = (pkg.CheckerboardAffineCoupling(C, D, S, float(cfg), hps) if kind == "ckbd"
               else pkg.ChannelwiseAffineCoupling(C, D, float(cfg), hps))
        st = O.random_state_from_shapes(O.coupling_state_shapes("", kind, C, D, R), seed=case["seed"])
        mod.load_state_dict(st, strict=True)
        mod = mod.to(DEV)
        mod.train(mode == "train")
        ost = {"c." + k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st.items()}
        ora = O.RealNVPOracle(ost, 3, 8, 8, R, 2, emulate_tf32=True)
        ora.training = mode == "train"
        xr = case["x"].clone().requires_grad_(True)
        y_o, J_o = ora.coupling("c", xr, kind=kind, cfg=cfg)
        x = case["x"].to(DEV).requires_grad_(True)
        if mode == "train":
            (y_o * case["gy"]).sum().add((J_o * case["gJ"]).sum()).backward()
            y, J = mod(x)
            (y * case["gy"].to(DEV)).sum().add((J * case["gJ"].to(DEV)).sum()).backward()
            named = dict(mod.named_parameters())
            gref = {k[2:]: v.grad for k, v in ost.items() if v.grad is not None}
            d, cs = grad_distance({k: named[k].grad for k in gref}, gref)
            worst[tag] = (rel(y, y_o), rel(J, J_o), rel(x.grad, xr.grad), d)
            print(tag, "train: y %.1e J %.1e gx %.1e grads rel-L2 %.1e cos %.5f" % (*worst[tag], cs))
            assert rel(y, y_o) < 1e-3 and rel(J, J_o) < 1e-3, (tag, rel(y, y_o), rel(J, J_o))
            assert rel(x.grad, xr.grad) < GATE_VJP and d < GATE_VJP, (tag, rel(x.grad, xr.grad), d, cs)
        else:
            with torch.no_grad():
                y, J = mod(x)
                xi, _ = mod(case["x"].to(DEV), reverse=True)
                ora_i = O.RealNVPOracle({"c." + k: v.clone() for k, v in st.items()}, 3, 8, 8, R, 2, emulate_tf32=True)
                ora_i.training = False
                xi_o, _ = ora_i.coupling("c", case["x"], reverse=True, kind=kind, cfg=cfg)
            worst[tag] = (rel(y, y_o), rel(J, J_o), rel(xi, xi_o))
            assert rel(y, y_o) < 1e-3 and rel(J, J_o) < 1e-3 and rel(xi, xi_o) < 2e-3, (tag, worst[tag])
    print(mode, {k: tuple(f"{e:.1e}" for e in v) for k, v in worst.items()})
    pkg.set_default_math("tf32")


# one coupling = 13 train-mode batch norms over 48-192 values at these fixture sizes: the ideal emulation itself is
# 1e-3 ... 9e-2 away from the fp32 VJP here (printed by oracle/make_golden_api.py's sibling probe)
GATE_VJP = 1e-1


def test_tf32_gradient_at_a_conditioned_point(pkg):
    """cfg A at batch 64 (BN statistics over >= 1024 values everywhere), mild scale: the tensor-core tier against the
    library's own fp32 tier, next to the distance an ideal TF32 implementation has from fp32 on the CPU."""
    ch, img, base, R, L, B, scale = 3, 64, 32, 4, 5, 64, 0.2
    cfg = (ch, img, base, R, L)
    c = dict(channels=ch, image=img, base_dim=base, res_blocks=R, num_scales=L)
    st0 = O.random_state(ch, img, base, R, L, seed=0, scale=scale)
    x_img = O.synthetic_images(B, ch, img, seed=0)
    x, _ = O.logit_forward(x_img, torch.rand(x_img.shape, generator=torch.Generator().manual_seed(1)))
    grads = {}
    lls = {}
    for mm in ("fp32", "tf32"):
        m = build(pkg, c, st0, mm)
        m.train()
        ll, ws = m(x.to(DEV))
        (-(ll).mean() + 5e-5 * ws).backward()
        grads[mm] = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters() if p.grad is not None}
        lls[mm] = ll.detach().cpu()
        del m
    d_cuda, cos_cuda = grad_distance(grads["tf32"], grads["fp32"])
    _, _, g_f = oracle_fwd_bwd(st0, cfg, x, emu=False)
    _, _, g_e = oracle_fwd_bwd(st0, cfg, x, emu=True)
    d_ideal, cos_ideal = grad_distance(g_e, g_f)
    d_fp32, _ = grad_distance(grads["fp32"], g_f)
    print(f"cfgA B=64: tf32-tier vs fp32-tier grad rel-L2 {d_cuda:.3f} cos {cos_cuda:.4f}; ideal TF32 vs fp32 (CPU) "
          f"{d_ideal:.3f} cos {cos_ideal:.4f}; fp32 tier vs fp32 oracle {d_fp32:.2e}; ll {rel(lls['tf32'], lls['fp32']):.2e}")
    assert rel(lls["tf32"], lls["fp32"]) < 1e-3
    assert d_fp32 < 2e-2
    assert d_cuda < 1.25 * d_ideal + 0.02 and cos_cuda > cos_ideal - 0.02, (d_cuda, d_ideal, cos_cuda, cos_ideal)


def _structured_images(n, seed):
    """Smooth synthetic 'images' (a few random blobs and gradients per image, uint8): learnable structure, no I/O."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 64), torch.linspace(-1, 1, 64), indexing="ij")
    out = torch.zeros(n, 3, 64, 64)
    for k in range(4):
        cx, cy = torch.rand(n, 3, 1, 1, generator=g) * 2 - 1, torch.rand(n, 3, 1, 1, generator=g) * 2 - 1
        sg = 0.15 + 0.5 * torch.rand(n, 3, 1, 1, generator=g)
        amp = torch.rand(n, 3, 1, 1, generator=g)
        out += amp * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg ** 2))
    out += 0.3 * torch.rand(n, 3, 1, 1, generator=g) * xx + 0.3 * torch.rand(n, 3, 1, 1, generator=g) * yy
    out = (out - out.amin((1, 2, 3), keepdim=True)) / (out.amax((1, 2, 3), keepdim=True) - out.amin((1, 2, 3), keepdim=True))
    return (out * 255).round().to(torch.uint8)


def test_tf32_training_tracks_fp32_training(pkg):
    """200 optimizer steps (train.py:176-200 semantics: logit, forward, loss, backward, Adam lr 5e-4 wd 5e-5) from the
    same seeded initialisation on the same structured synthetic data, once per tier.  The per-step gradients differ by
    the TF32 operand noise, so the two trajectories decorrelate like two runs with different summation orders would;
    what must hold is that the tier TRAINS the same.  At step 200 the loss still falls by ~2 % per 10 steps, so the
    gates are expressed against that slope: the TF32 run is at most 20 steps behind the fp32 run at the end, the bits/dim
    level reached (train.py:203-207 formula) agrees within 3 % and the windowed curves within 6 % all along (measured on
    B200 over several runs -- the fp32 tier's atomics make every run slightly different: 0.2 % ... 1.4 % and
    2.6 % ... 2.9 %)."""
    B, steps, D = 128, 200, 64 * 64 * 3
    data = _structured_images(1024, seed=3).to(DEV)
    curves = {}
    for mm in ("fp32", "tf32"):
        torch.manual_seed(999)                                   # main.py:58-59 default seed
        prior = torch.distributions.Normal(torch.tensor(0., device=DEV), torch.tensor(1., device=DEV), validate_args=False)
        m = pkg.RealNVP(3, 64, prior, pkg.Hyperparameters(32, 4, True, True, True, True)).to(DEV)
        m.set_math(mm)
        m.train()
        opt = pkg.rnvp_optim.Adam(m, lr=5e-4, weight_decay=5e-5)
        g = torch.Generator().manual_seed(11)
        torch.manual_seed(5)                                     # dequantisation noise stream
        bpd = []
        for it in range(steps):
            idx = torch.randint(0, data.shape[0], (B,), generator=g).to(DEV)
            opt.zero_grad()
            x, logdet = pkg.logit_transform(data[idx])
            ll, ws = m(x)
            logll = (ll + logdet).mean()
            (-logll + 5e-5 * ws).backward()
            opt.step()
            bpd.append((-float(logll) + math.log(256.0) * D) / (D * math.log(2.0)))
        curves[mm] = torch.tensor(bpd)
        del m, opt
        torch.cuda.empty_cache()
    f, t = curves["fp32"], curves["tf32"]
    # compare the smoothed curves (window 10) after the first steps and the final level
    w = 10
    fs, ts = f.unfold(0, w, w).mean(1), t.unfold(0, w, w).mean(1)
    dev = ((ts - fs).abs() / fs).max()
    final = abs(float(ts[-5:].mean() - fs[-5:].mean())) / float(fs[-5:].mean())
    print(f"bits/dim fp32 tier {fs[0]:.3f} -> {fs[-1]:.3f}; tf32 tier {ts[0]:.3f} -> {ts[-1]:.3f}; max windowed deviation "
          f"{dev:.2e}; final level (last 50 steps) differs by {final:.2e}")
    assert fs[-1] < fs[0] - 0.5 and ts[-1] < ts[0] - 0.5, "the model did not train"
    assert float(ts[-1]) <= float(fs[-3]), ("the TF32 run lags the fp32 run by more than 20 steps", fs, ts)
    assert final < 3e-2, (final, fs, ts)
    assert dev < 6e-2, (dev, fs, ts)
