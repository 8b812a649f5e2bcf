"""Print parity numbers of the CUDA path vs the CPU oracle (fp32 and tf32 tiers). GPU box only."""
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


def run(channels, image, base, R, L, B, scale, seed=0):
    st0 = O.random_state(channels, image, base, R, L, seed=seed, scale=scale)
    x_img = O.synthetic_images(B, channels, image, seed=seed)
    g = torch.Generator().manual_seed(1)
    x, ld_logit = O.logit_forward(x_img, torch.rand(x_img.shape, generator=g))
    ost = {k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st0.items()}
    ora = O.RealNVPOracle(ost, channels, image, base, R, L)
    t0 = time.time()
    z, ld, lp = ora.log_prob_parts(x)
    ll = lp + ld
    (-(ll + ld_logit).mean() + 5e-5 * ora.weight_scale()).backward()
    print(f"  oracle fwd+bwd {time.time()-t0:.1f}s  ll={ll.detach()[:2].tolist()}")
    gref = {k: v.grad for k, v in ost.items() if v.grad is not None}
    prior = torch.distributions.Normal(torch.tensor(0., device=DEV), torch.tensor(1., device=DEV))
    for math in ("fp32", "tf32"):
        m = pkg.RealNVP(channels, image, prior, pkg.Hyperparameters(base, R, True, True, True, True), num_scales=L)
        m.load_state_dict(st0); m = m.to(DEV); m.set_math(math); m.train()
        lld, ws = m(x.to(DEV))
        (-(lld + ld_logit.to(DEV)).mean() + 5e-5 * ws).backward()
        num = den = dot = n1 = 0.0
        for k, p in m.named_parameters():
            if p.grad is None: continue
            a, b = p.grad.cpu().double().flatten(), gref[k].double().flatten()
            num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum()); dot += float((a * b).sum()); n1 += float((a ** 2).sum())
        m2 = pkg.RealNVP(channels, image, prior, pkg.Hyperparameters(base, R, True, True, True, True), num_scales=L)
        m2.load_state_dict(st0); m2 = m2.to(DEV); m2.set_math(math); m2.train()
        zd, ldd, _ = m2.latent(x.to(DEV))
        print(f"  [{math}] train: ll rel {rel(lld, ll):.2e}  logdet rel {rel(ldd, ld):.2e}  z maxabs {float((zd.cpu()-z.detach()).abs().max()):.2e}"
              f"  grad rel-L2 {(num/den)**0.5:.2e} cos {dot/(n1*den)**0.5:.6f}")
        # converged running stats, eval
        with torch.no_grad():
            for _ in range(30): m2(x.to(DEV))
        sd = {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()}
        oe = O.RealNVPOracle(sd, channels, image, base, R, L); oe.training = False
        m2.eval()
        with torch.no_grad():
            z_e, ld_e, lp_e = oe.log_prob_parts(x)
            zz, ldd, lld = m2.latent(x.to(DEV))
            rec = m2.g(zz)
            rec_o = oe.g(z_e)
        print(f"  [{math}] eval(converged): ll rel {rel(lld, lp_e+ld_e):.2e} logdet rel {rel(ldd, ld_e):.2e}  recon {float((rec-x.to(DEV)).abs().max()):.2e} (oracle {float((rec_o-x).abs().max()):.2e})")


for name, args in [("tiny 32px base4 R2 B4 scale .7", (3, 32, 4, 2, 5, 4, 0.7, 3)),
                   ("cfgA B=8 scale .7", (3, 64, 32, 4, 5, 8, 0.7)),
                   ("cfgA B=8 scale .2", (3, 64, 32, 4, 5, 8, 0.2))]:
    print(name)
    run(*args)
