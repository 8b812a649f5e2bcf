"""The CPU oracle against the fixtures produced by the reference itself (oracle/make_golden.py)."""
import hashlib
import os

import pytest
import torch

import realnvp_oracle as O
from _util import compare_grads


def sha(state):
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def clone(state, grad=False):
    out = {}
    for k, v in state.items():
        t = v.detach().clone()
        if grad and O.is_trainable(k) and t.is_floating_point():
            t.requires_grad_(True)
        out[k] = t
    return out


def test_logit_and_layout(golden_dir):
    fix = torch.load(os.path.join(golden_dir, "misc.pt"))
    lg = fix["logit"]
    y, ld = O.logit_forward(lg["x"], lg["noise"])
    assert torch.equal(y, lg["y"])
    assert torch.allclose(ld, lg["logdet"], rtol=1e-6)
    assert torch.allclose(O.logit_inverse(lg["y"]), lg["inv"], rtol=1e-6, atol=1e-7)
    lay = fix["layout"]
    assert torch.equal(O.squeeze(lay["t"]), lay["squeeze"])
    assert torch.equal(O.undo_squeeze(lay["squeeze"]), lay["t"])
    on, off = O.factor_out(lay["t"])
    assert torch.equal(on, lay["on"]) and torch.equal(off, lay["off"])
    assert torch.equal(O.restore(on, off), lay["t"])


def test_logit_edge_cases():
    # empty batch and the extreme pixel values (0 and 255) stay finite
    y, ld = O.logit_forward(torch.zeros(0, 3, 4, 4), torch.zeros(0, 3, 4, 4))
    assert y.shape == (0, 3, 4, 4) and ld.shape == (0,)
    x = torch.tensor([0.0, 1.0]).reshape(2, 1, 1, 1)
    y, ld = O.logit_forward(x, torch.tensor([0.0, 0.999999]).reshape(2, 1, 1, 1))
    assert torch.isfinite(y).all() and torch.isfinite(ld).all()


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_couplings(golden_dir, mode):
    fix = torch.load(os.path.join(golden_dir, "couplings.pt"))
    for tag, case in fix.items():
        kind, C, D, cfg, R = case["kind"], case["C"], case["D"], case["cfg"], case["R"]
        st = O.random_state_from_shapes(O.coupling_state_shapes("", kind, C, D, R), seed=case["seed"])
        assert sha(st) == case["state_sha256"], tag
        ost = clone({"c." + k: v for k, v in st.items()}, grad=True)
        ora = O.RealNVPOracle(ost, 3, 8, 8, R, 2)
        ora.training = mode == "train"
        x = case["x"].clone().requires_grad_(True)
        y, J = ora.coupling("c", x, kind=kind, cfg=cfg)
        (y * case["gy"]).sum().add((J * case["gJ"]).sum()).backward()
        ref = case[mode]
        assert rel(y.detach(), ref["y"]) < 1e-6 and rel(J.detach(), ref["J"]) < 1e-6, tag
        assert rel(x.grad, ref["gx"]) < 1e-4, tag
        for k, g in ref["grads"].items():
            assert rel(ost["c." + k].grad, g) < 1e-3 or float(g.abs().max()) < 1e-5, (tag, k)
        ora2 = O.RealNVPOracle(clone({"c." + k: v for k, v in st.items()}), 3, 8, 8, R, 2)
        ora2.training = mode == "train"
        with torch.no_grad():
            xi, _ = ora2.coupling("c", case["x"], reverse=True, kind=kind, cfg=cfg)
        assert rel(xi, ref["inv"]) < 1e-5, tag


@pytest.mark.parametrize("name", ["tiny_32px_b4", "small_64px_b2"])
def test_full_model(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["config"]
    st0 = O.random_state(c["channels"], c["image"], c["base_dim"], c["res_blocks"], c["num_scales"], seed=c["seed"])
    assert sha(st0) == fix["state_sha256"]
    x = fix["x"]
    st = clone(st0, grad=True)
    ora = O.RealNVPOracle(st, c["channels"], c["image"], c["base_dim"], c["res_blocks"], c["num_scales"])
    ll, ws = ora.forward(x)
    assert rel(ll.detach(), fix["train_ll"]) < 1e-6
    assert rel(ws.detach(), fix["train_ws"]) < 1e-6
    (-(ll).mean() + 5e-5 * ws).backward()        # the constant logit log-det does not change gradients
    grel, worst, wk = compare_grads({k: st[k].grad for k in fix["train_grads"]}, fix["train_grads"],
                                    fix["train_grad_norms"])
    assert grel < 1e-3 and worst < 2e-2, (grel, worst, wk)
    for k, v in fix["state_after_train"].items():
        assert torch.allclose(st[k].detach().to(v.dtype), v, rtol=1e-4, atol=1e-5), k
    ora2 = O.RealNVPOracle(clone(st0), c["channels"], c["image"], c["base_dim"], c["res_blocks"], c["num_scales"])
    with torch.no_grad():
        z, J = ora2.f(x)
    assert rel(z, fix["train_z"]) < 1e-5 and rel(J, fix["train_J"]) < 1e-5
    ora3 = O.RealNVPOracle(clone(st0), c["channels"], c["image"], c["base_dim"], c["res_blocks"], c["num_scales"])
    ora3.training = False
    with torch.no_grad():
        assert rel(ora3.log_prob(x), fix["eval_ll"]) < 1e-5
        assert rel(ora3.g(fix["z_sample"]), fix["eval_g"]) < 1e-5
        # round trip at the reference's own floor
        rec = ora3.g(ora3.f(x)[0])
    assert float((rec - x).abs().max()) < max(10 * fix["eval_recon_err_ref"], 1e-4)


def test_two_scale_variant_runs():
    """BASELINE config 3 interpretation: the same class truncated to two scales."""
    st = O.random_state(3, 8, 4, 1, 2, seed=1)
    ora = O.RealNVPOracle(st, 3, 8, 4, 1, 2)
    assert len(ora.specs) == 10
    ora.training = False
    x = torch.randn(2, 3, 8, 8)
    with torch.no_grad():
        z, J = ora.f(x)
        assert z.shape == x.shape and J.shape == x.shape
        assert float((ora.g(z) - x).abs().max()) < 1e-3
