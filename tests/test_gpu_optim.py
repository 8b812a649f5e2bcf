"""-m gpu: the fused Adam step (rnvp_adam_step) against torch.optim.Adam on the same gradients, and the
torch-format optimizer checkpoint in both directions (train.py:134, 150, 200, 250)."""
import pytest
import torch

import realnvp_oracle as O
from _util import rel

pytestmark = pytest.mark.gpu
DEV = "cuda"
CFG = dict(channels=3, image=32, base_dim=32, res_blocks=1, num_scales=3)


def _model(pkg, state):
    prior = torch.distributions.Normal(torch.tensor(0., device=DEV), torch.tensor(1., device=DEV), validate_args=False)
    m = pkg.RealNVP(CFG["channels"], CFG["image"], prior, pkg.Hyperparameters(CFG["base_dim"], CFG["res_blocks"], True, True, True, True),
                    num_scales=CFG["num_scales"])
    m.load_state_dict(state, strict=True)
    m = m.to(DEV)
    m.set_math("fp32")
    return m.train()


def _step(m, opt, x):
    opt.zero_grad()
    ll, ws = m(x)
    loss = -ll.mean() + 5e-5 * ws
    loss.backward()
    opt.step()
    return float(loss.detach())


def _paired_step(a, oa, b, ob, x):
    """Both optimizers see IDENTICAL gradients (those of model b): Adam turns the noise-level gradients of the
    analytically gradient-free biases into +-lr updates, so two independent backward passes
    would not agree on them."""
    ob.zero_grad()
    ll, ws = b(x)
    (-ll.mean() + 5e-5 * ws).backward()
    pb = dict(b.named_parameters())
    for k, p in a.named_parameters():
        p.grad = None if pb[k].grad is None else pb[k].grad.detach().clone()
    oa.step()
    ob.step()


def test_fused_adam_matches_torch_adam(pkg):
    st = O.random_state(CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"], seed=3)
    a, b = _model(pkg, st), _model(pkg, st)
    kw = dict(lr=5e-4, weight_decay=5e-5)                        # train.py:134
    oa = torch.optim.Adam(a.parameters(), **kw)
    ob = pkg.rnvp_optim.Adam(b, **kw)
    x = O.logit_forward(O.synthetic_images(4, 3, CFG["image"], seed=1), torch.rand(4, 3, CFG["image"], CFG["image"]))[0].to(DEV)
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for i in range(3):
        _paired_step(a, oa, b, ob, x)
        # keep the two replicas on the same trajectory: only the optimizer arithmetic is under test
        worst = max(rel(pb[k], pa[k]) for k in pa)
        assert worst < 1e-5, (i, worst)
    # gradients were cleared by the fused step; the views stay installed
    assert all(float(p.grad.abs().max()) == 0.0 for p in b.parameters() if p.grad is not None)
    # optimizer checkpoint has torch's layout and moves both ways
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0]["params"] == sb["param_groups"][0]["params"]
    assert set(sa["state"]) == set(sb["state"])
    for k in sa["state"]:
        assert float(sb["state"][k]["step"]) == float(sa["state"][k]["step"]) == 3.0
        assert rel(sb["state"][k]["exp_avg"], sa["state"][k]["exp_avg"]) < 1e-5
        assert rel(sb["state"][k]["exp_avg_sq"], sa["state"][k]["exp_avg_sq"]) < 1e-4     # fp32 rounding order of g + wd*p, squared
    oa2 = torch.optim.Adam(a.parameters(), **kw)
    oa2.load_state_dict(sb)                                       # fused -> torch
    ob2 = pkg.rnvp_optim.Adam(b, **kw)
    ob2.load_state_dict(sa)                                       # torch -> fused
    _paired_step(a, oa2, b, ob2, x)
    worst = max(rel(pb[k], pa[k]) for k in pa)
    assert worst < 2e-5, worst
    assert float(ob2.state_dict()["state"][next(iter(sa["state"]))]["step"]) == 4.0


def test_zero_grad_after_discarded_backward(pkg):
    """A backward whose gradients are thrown away (no step) must still be cleared by zero_grad."""
    st = O.random_state(CFG["channels"], CFG["image"], CFG["base_dim"], CFG["res_blocks"], CFG["num_scales"], seed=4)
    m = _model(pkg, st)
    opt = pkg.rnvp_optim.Adam(m, lr=5e-4)
    x = torch.randn(2, 3, CFG["image"], CFG["image"], device=DEV)
    _step(m, opt, x)
    ll, ws = m(x)
    (-ll.mean()).backward()
    assert any(float(p.grad.abs().max()) > 0 for p in m.parameters() if p.grad is not None)
    opt.zero_grad()
    assert all(float(p.grad.abs().max()) == 0.0 for p in m.parameters() if p.grad is not None)
