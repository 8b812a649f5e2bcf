"""Print parity numbers of the CUDA path vs the CPU oracle (fp32 and tf32 tiers). GPU box only.

For every configuration three comparisons are printed:
  fp32 tier  vs fp32 oracle            -- the 1e-5 tier
  tf32 tier  vs fp32 oracle            -- what TF32 operands cost (inherent to ANY TF32 implementation)
  tf32 tier  vs TF32-EMULATING oracle  -- kernel correctness of the tensor-core tier (rounding points mirrored)
and `emu vs fp32`, the same inherent cost measured on the CPU with ideal round-to-nearest operands.
"""
import importlib, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import realnvp_oracle as O
pkg = importlib.import_module("dl-normalizing-flows_b200")
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max())


def gcmp(got, ref):
    num = den = dot = n1 = 0.0
    for k, b in ref.items():
        a, b = got[k].detach().cpu().double().flatten(), b.double().flatten()
        num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum()); dot += float((a * b).sum()); n1 += float((a ** 2).sum())
    return (num / den) ** 0.5, dot / (n1 * den) ** 0.5


def oracle_run(st0, cfg, x, ld_logit, emu):
    ost = {k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st0.items()}
    ora = O.RealNVPOracle(ost, *cfg, emulate_tf32=emu)
    z, ld, lp = ora.log_prob_parts(x)
    ll = lp + ld
    (-(ll + ld_logit).mean() + 5e-5 * ora.weight_scale()).backward()
    return ll.detach(), ld.detach(), z.detach(), {k: v.grad for k, v in ost.items() if v.grad is not None}


def run(channels, image, base, R, L, B, scale, seed=0):
    cfg = (channels, image, base, R, L)
    st0 = O.random_state(channels, image, base, R, L, seed=seed, scale=scale)
    x_img = O.synthetic_images(B, channels, image, seed=seed)
    g = torch.Generator().manual_seed(1)
    x, ld_logit = O.logit_forward(x_img, torch.rand(x_img.shape, generator=g))
    t0 = time.time()
    ref = {emu: oracle_run(st0, cfg, x, ld_logit, emu) for emu in (False, True)}
    r2, cs = gcmp(ref[True][3], ref[False][3])
    print(f"  oracles {time.time()-t0:.1f}s | emu vs fp32 (CPU, ideal TF32): ll {rel(ref[True][0], ref[False][0]):.2e} "
          f"logdet {rel(ref[True][1], ref[False][1]):.2e} grad rel-L2 {r2:.2e} cos {cs:.6f}")
    prior = torch.distributions.Normal(torch.tensor(0., device=DEV), torch.tensor(1., device=DEV))
    hps = pkg.Hyperparameters(base, R, True, True, True, True)
    grads = {}
    for math in ("fp32", "tf32"):
        m = pkg.RealNVP(channels, image, prior, hps, num_scales=L)
        m.load_state_dict(st0); m = m.to(DEV); m.set_math(math); m.train()
        lld, ws = m(x.to(DEV))
        (-(lld + ld_logit.to(DEV)).mean() + 5e-5 * ws).backward()
        grads[math] = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters() if p.grad is not None}
        m2 = pkg.RealNVP(channels, image, prior, hps, num_scales=L)
        m2.load_state_dict(st0); m2 = m2.to(DEV); m2.set_math(math); m2.train()
        zd, ldd, _ = m2.latent(x.to(DEV))
        for tag, emu in (("fp32-oracle", False), ("emu-oracle ", True)):
            if math == "fp32" and emu:
                continue
            ll, ld, z, gref = ref[emu]
            r2, cs = gcmp(grads[math], gref)
            print(f"  [{math}] train vs {tag}: ll {rel(lld, ll):.2e} logdet {rel(ldd, ld):.2e} z maxabs "
                  f"{float((zd.cpu()-z).abs().max()):.2e} grad rel-L2 {r2:.2e} cos {cs:.6f}")
        # converged running stats, eval
        with torch.no_grad():
            for _ in range(30): m2(x.to(DEV))
        sd = {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()}
        m2.eval()
        with torch.no_grad():
            zz, ldd, lld = m2.latent(x.to(DEV))
            rec = m2.g(zz)
        for tag, emu in (("fp32-oracle", False), ("emu-oracle ", True)):
            if math == "fp32" and emu:
                continue
            oe = O.RealNVPOracle({k: v.clone() for k, v in sd.items()}, *cfg, emulate_tf32=emu); oe.training = False
            with torch.no_grad():
                z_e, ld_e, lp_e = oe.log_prob_parts(x)
                rec_o = oe.g(z_e)
            print(f"  [{math}] eval(converged) vs {tag}: ll {rel(lld, lp_e+ld_e):.2e} logdet {rel(ldd, ld_e):.2e}  recon "
                  f"{float((rec-x.to(DEV)).abs().max()):.2e} (oracle {float((rec_o-x).abs().max()):.2e})")
    r2, cs = gcmp(grads["tf32"], grads["fp32"])
    print(f"  tf32 tier vs fp32 tier (both CUDA): grad rel-L2 {r2:.2e} cos {cs:.6f}")


cases = [("tiny 32px base4 R2 B4 scale .7", (3, 32, 4, 2, 5, 4, 0.7, 3)),
         ("cfgA B=8 scale .7", (3, 64, 32, 4, 5, 8, 0.7)),
         ("cfgA B=8 scale .2", (3, 64, 32, 4, 5, 8, 0.2))]
if "--big" in sys.argv:
    cases.append(("cfgA B=64 scale .2", (3, 64, 32, 4, 5, 64, 0.2)))
for name, args in cases:
    print(name)
    run(*args)
