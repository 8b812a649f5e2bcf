"""Generate tests/golden/api.pt from the UNMODIFIED reference  --  build-container only (needs /root/reference).

Fixtures for the API-completeness rows, all outputs of the reference itself:
  * ``coupling(x, reverse=True)`` -> (x, log_rescale) for the states of tests/golden/couplings.pt;
  * stand-alone ``WeightNormConv2d`` / ``ResidualBlock`` / ``ResidualModule`` forwards and input gradients under
    seeded initialisation (the drop-in's initialisation is bit-identical, tests/test_boundary_cpu.py);
  * the hyper-parameter branches train.py never selects (bottleneck / skip / weight_norm / coupling_bn = False,
    res_blocks = 0): train-mode log-likelihood, weight_scale, a few gradients, eval-mode g().
Run:  python oracle/make_golden_api.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import realnvp_oracle as O          # noqa: E402
from make_golden import ROOT, clone_state, load_reference          # noqa: E402

HPS_CASES = [(False, True, True, True, 1), (True, False, True, True, 1), (True, True, False, True, 1),
             (True, True, True, False, 1), (True, True, True, True, 0), (False, False, False, False, 0)]


def main():
    ref_flow, ref_mod, ref_utils = load_reference()
    torch.set_num_threads(8)
    fix = {}
    # ---- reverse=True returns (x, log_rescale) -------------------------------------------------------------
    cpl = torch.load(os.path.join(ROOT, "tests", "golden", "couplings.pt"))
    rev = {}
    for tag, case in cpl.items():
        kind, C, S, D, cfg, R = case["kind"], case["C"], case["S"], case["D"], case["cfg"], case["R"]
        hps = ref_utils.Hyperparameters(8, R, True, True, True, True)
        mod = (ref_mod.CheckerboardAffineCoupling(C, D, S, float(cfg), hps) if kind == "ckbd"
               else ref_mod.ChannelwiseAffineCoupling(C, D, float(cfg), hps))
        st = O.random_state_from_shapes(O.coupling_state_shapes("", kind, C, D, R), seed=case["seed"])
        mod.load_state_dict(clone_state(st), strict=True)
        mod.eval()
        with torch.no_grad():
            xi, lj = mod(case["x"], reverse=True)
        assert torch.equal(xi, case["eval"]["inv"])
        rev[tag] = {"inv": xi.clone(), "inv_logJ": lj.clone()}
    fix["reverse"] = rev
    # ---- stand-alone helper modules -------------------------------------------------------------------------
    g = torch.Generator().manual_seed(21)
    sub = {}
    for name, make, shape in (
            ("WeightNormConv2d", lambda: ref_mod.WeightNormConv2d(12, 24, (3, 3), 1, 1, True, True, True), (3, 12, 8, 8)),
            ("ResidualBlock", lambda: ref_mod.ResidualBlock(32, True, True), (4, 32, 8, 8)),
            ("ResidualModule", lambda: ref_mod.ResidualModule(13, 32, 12, 2, True, True, True), (4, 13, 16, 16))):
        torch.manual_seed(5)
        m = make()
        m.train()
        x = torch.randn(*shape, generator=g).requires_grad_(True)
        gy_shape = m(x.detach()).shape                      # (also advances the running statistics once)
        torch.manual_seed(5)
        m = make()
        m.train()
        y = m(x)
        gy = torch.randn(*gy_shape, generator=g)
        (y * gy).sum().backward()
        sub[name] = {"seed": 5, "x": x.detach().clone(), "gy": gy, "y": y.detach().clone(), "gx": x.grad.clone()}
    fix["submodules"] = sub
    # ---- non-default hyper-parameter branches --------------------------------------------------------------------
    prior = torch.distributions.Normal(torch.tensor(0.), torch.tensor(1.), validate_args=False)
    x = torch.randn(3, 3, 32, 32, generator=g)
    zs = torch.randn(2, 3, 32, 32, generator=g)
    hp = {}
    for bott, skip, wn, cbn, R in HPS_CASES:
        torch.manual_seed(3)
        m = ref_flow.RealNVP(3, 32, prior, ref_utils.Hyperparameters(8, R, bott, skip, wn, cbn))
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith(".scale"):
                    p.fill_(0.5)
        m.train()
        ll, ws = m(x)
        loss = -ll.mean() + (5e-5 * ws if ws != 0 else 0.0)
        loss.backward()
        # fixtures stay small: the scalar / per-channel gradients of every coupling and leading slices of the last
        # coupling's tensors
        grads = {}
        for n, p in m.named_parameters():
            if p.grad is None:
                continue
            if n.endswith(("scale", "scale_shift", "in_bn.weight", "in_bn.bias")):
                grads[n] = p.grad.clone()
            elif "s5_ckbd.3" in n:
                grads[n] = p.grad.flatten()[:64].clone()
        m.eval()
        with torch.no_grad():
            xs = m.g(zs)
            lle, _ = m(x)
        hp[(bott, skip, wn, cbn, R)] = {"train_ll": ll.detach().clone(), "train_ws": torch.as_tensor(float(ws)),
                                       "grads": grads, "eval_g": xs.clone(), "eval_ll": lle.clone()}
    fix["hps"] = {"seed": 3, "scale": 0.5, "x": x, "z": zs, "cases": hp}
    torch.save(fix, os.path.join(ROOT, "tests", "golden", "api.pt"))
    print("written tests/golden/api.pt")


if __name__ == "__main__" and "--checkpoint" not in sys.argv:
    main()


def checkpoint_case(quiet=False):
    """A checkpoint pair written by the REFERENCE's own calls (train.py:249-250: torch.save(model.state_dict()),
    torch.save(optimizer.state_dict())) after two Adam steps of its loop, plus what its third step yields after a
    resume (train.py:139-154) -- the interop fixture for SURVEY.md 8f-3."""
    ref_flow, ref_mod, ref_utils = load_reference()
    # 14 MB of reference-written checkpoint: a build output next to the byte-compiled reference (git-ignored, shipped
    # to the GPU box with the snapshot), regenerated by __graft_entry__.build() wherever /root/reference is mounted
    out = os.path.join(HERE, "_ref", "ckpt")
    os.makedirs(out, exist_ok=True)
    prior = torch.distributions.Normal(torch.tensor(0.), torch.tensor(1.), validate_args=False)

    def make():
        torch.manual_seed(999)
        m = ref_flow.RealNVP(3, 32, prior, ref_utils.Hyperparameters(4, 1, True, True, True, True))
        return m, torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=5e-5)

    def batch(i):
        x_img = O.synthetic_images(6, 3, 32, seed=40 + i)
        noise = torch.rand(x_img.shape, generator=torch.Generator().manual_seed(90 + i))
        return O.logit_forward(x_img, noise)            # == utils.logit_transform with this noise (tests pin that)

    def step(m, opt, i):
        x, logdet = batch(i)
        opt.zero_grad()
        logll, ws = m(x)
        logll = (logll + logdet).mean()
        (-logll + 5e-5 * ws).backward()
        opt.step()
        return float(logll)

    m, opt = make()
    m.train()
    l0, l1 = step(m, opt, 0), step(m, opt, 1)
    torch.save(m.state_dict(), os.path.join(out, "realnvp_state.pt"))
    torch.save(opt.state_dict(), os.path.join(out, "realnvp_state_optim.pt"))
    m2, opt2 = make()
    m2.load_state_dict(torch.load(os.path.join(out, "realnvp_state.pt")))
    opt2.load_state_dict(torch.load(os.path.join(out, "realnvp_state_optim.pt")))
    m2.train()
    l2 = step(m2, opt2, 2)
    key = "s3_chan.1.block.1.out_block.2.conv.weight_v"
    torch.save({"logll": [l0, l1, l2], "batches": [batch(i) for i in range(3)],
                "after_step3": {key: m2.state_dict()[key].clone(), "s1_ckbd.0.scale": m2.state_dict()["s1_ckbd.0.scale"].clone()}},
               os.path.join(out, "expect.pt"))
    if not quiet:
        print("written", out, [os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)])


if __name__ == "__main__" and "--checkpoint" in sys.argv:
    checkpoint_case()
