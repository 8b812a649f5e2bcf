"""Host-side checks that run without a GPU: the C-ABI library loads and exports every symbol the
header declares, the plan's parameter table mirrors the reference's state dict, and the drop-in
modules are state-dict / parameter-order / initialisation compatible with the reference."""
import ctypes as C
import os
import re

import pytest
import torch

import realnvp_oracle as O
from _util import ROOT, sha


def test_header_symbols_exported(pkg):
    header = open(os.path.join(ROOT, "include", "rnvp.h")).read()
    declared = sorted(set(re.findall(r"\b(rnvp_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    lib = pkg.rnvp_cabi.lib
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rnvp.h but not exported"
    assert sorted(pkg.rnvp_cabi.EXPORTS) == declared
    assert b"sm_100a" in lib.rnvp_version()


def test_plan_table_matches_state_dict(pkg):
    cabi = pkg.rnvp_cabi
    cfg = cabi.Config(3, 64, 32, 4, 5, 0.0, 1.0)
    h = C.c_void_p()
    cabi.check(cabi.lib.rnvp_plan_create(C.byref(cfg), C.byref(h)))
    try:
        n = cabi.lib.rnvp_plan_num_couplings(h)
        slots = cabi.lib.rnvp_plan_slots_per_coupling(h)
        assert n == 28 and slots == 8 + 3 * 19 + 4 * 13
        specs = O.coupling_specs(3, 64, 32, 5)
        shapes = O.state_shapes(3, 64, 32, 4, 5)
        seen = set()
        for i, (name, kind, Cc, S, D, mcfg) in enumerate(specs):
            buf = C.create_string_buffer(64)
            vals = [C.c_int() for _ in range(5)]
            cabi.check(cabi.lib.rnvp_plan_coupling_info(h, i, buf, 64, *[C.byref(v) for v in vals]))
            assert buf.value.decode() == name
            assert [v.value for v in vals] == [0 if kind == "ckbd" else 1, Cc, S, D, mcfg]
            for s in range(slots):
                sn = cabi.lib.rnvp_plan_slot_name(h, s)
                if sn:
                    key = name + "." + sn.decode()
                    assert key in shapes, key
                    seen.add(key)
        missing = [k for k in shapes if k not in seen and not k.endswith("num_batches_tracked")]
        assert not missing, missing[:5]
        # workspace query is pure host arithmetic
        w_inf = cabi.lib.rnvp_plan_workspace_bytes(h, 64, 0)
        w_trn = cabi.lib.rnvp_plan_workspace_bytes(h, 64, 1)
        assert 0 < w_inf < w_trn
    finally:
        cabi.lib.rnvp_plan_destroy(h)


def test_math_tiers_and_workspace_sizes(pkg):
    """Host arithmetic of the three arithmetic tiers: the 3xTF32 tier keeps a lo copy of every weight layout (the
    weight arena doubles), workspace mode 2 (keeps activations) is larger than the lean mode 1, and an unknown tier
    is rejected."""
    cabi = pkg.rnvp_cabi
    assert cabi.MATH_BY_NAME == {"fp32": 0, "tf32": 1, "tf32x3": 2}
    cfg = cabi.Config(3, 64, 32, 4, 5, 0.0, 1.0)
    h = C.c_void_p()
    cabi.check(cabi.lib.rnvp_plan_create(C.byref(cfg), C.byref(h)))
    try:
        size = {}
        for name, mode in cabi.MATH_BY_NAME.items():
            cabi.check(cabi.lib.rnvp_plan_set_math(h, mode))
            size[name] = [cabi.lib.rnvp_plan_workspace_bytes(h, 8, m) for m in (0, 1, 2)]
            assert 0 < size[name][0] < size[name][1] <= size[name][2], (name, size[name])
        # 120.15 M parameters -> two padded layouts of every conv (forward + dgrad operand) ~ 1 GB of fp32 weights;
        # the lo copy of the 3xTF32 tier adds the same amount again, in every workspace mode
        extra = [a - b for a, b in zip(size["tf32x3"], size["tf32"])]
        assert all(0.9e9 < e < 1.2e9 for e in extra), extra
        assert cabi.lib.rnvp_plan_set_math(h, 7) != 0
        assert b"math" in cabi.lib.rnvp_last_error()
    finally:
        cabi.lib.rnvp_plan_destroy(h)


def test_plan_rejects_bad_config(pkg):
    cabi = pkg.rnvp_cabi
    h = C.c_void_p()
    for cfg in (cabi.Config(3, 64, 32, 0, 5, 0.0, 1.0),      # res_blocks = 0 uses another topology
                cabi.Config(3, 60, 32, 4, 5, 0.0, 1.0),      # 60 not divisible by 16
                cabi.Config(3, 64, 30, 4, 5, 0.0, 1.0)):     # base_dim not a multiple of 4
        assert cabi.lib.rnvp_plan_create(C.byref(cfg), C.byref(h)) == -1
        assert cabi.lib.rnvp_last_error()
    with pytest.raises(cabi.RnvpError):
        cabi.check(cabi.lib.rnvp_plan_create(C.byref(cabi.Config(3, 64, 32, 0, 5, 0.0, 1.0)), C.byref(h)))


def _model(pkg, channels=3, image=32, base=4, R=2, scales=5):
    prior = torch.distributions.Normal(torch.tensor(0.), torch.tensor(1.), validate_args=False)
    return pkg.RealNVP(channels, image, prior, pkg.Hyperparameters(base, R, True, True, True, True),
                       **({} if scales == 5 else {"num_scales": scales}))


def test_state_dict_and_param_order(pkg, golden_dir):
    m = _model(pkg)
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == list(O.state_shapes(3, 32, 4, 2, 5).items())
    fix = torch.load(os.path.join(golden_dir, "misc.pt"))
    assert [(n, bool(p.requires_grad)) for n, p in m.named_parameters()] == fix["param_names_3x32_b4_r2"]
    # checkpoints interchange: a reference-layout state loads strictly and round-trips
    st = O.random_state(3, 32, 4, 2, 5, seed=3)
    m.load_state_dict(st, strict=True)
    assert sha(m.state_dict()) == sha(st)
    # optimizer state is index based: same parameter count and order
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=5e-5)
    assert len(opt.param_groups[0]["params"]) == len(fix["param_names_3x32_b4_r2"])


def test_init_is_bit_identical_to_reference(pkg, golden_dir):
    fix = torch.load(os.path.join(golden_dir, "misc.pt"))
    torch.manual_seed(999)
    m = _model(pkg)
    assert sha(m.state_dict()) == fix["init_sha256_seed999_3x32_b4_r2"]


def test_cfg_a_inventory(pkg):
    m = _model(pkg, 3, 64, 32, 4)
    params = list(m.parameters())
    assert len(m.state_dict()) == 3472 and len(params) == 2212
    assert sum(p.requires_grad for p in params) == 1960
    assert sum(p.numel() for p in params if p.requires_grad) == 120_089_100 or \
        abs(sum(p.numel() for p in params if p.requires_grad) - 120.09e6) < 0.01e6


def test_two_scale_constructor(pkg):
    m = _model(pkg, 3, 32, 64, 1, scales=2)
    assert [n for n, _ in m.named_children()] == ["s1_ckbd", "s1_chan", "s2_ckbd"]
    assert len(m._couplings()) == 10


def test_no_cpu_fallback(pkg):
    m = _model(pkg)
    x = torch.zeros(2, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.g(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.s1_ckbd[0](x)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            pkg.logit_transform(x)


def test_fused_adam_host_side(pkg):
    """rnvp_optim.Adam is a torch Optimizer with torch.optim.Adam's group layout; on a CPU model its step
    raises instead of falling back."""
    m = _model(pkg)
    opt = pkg.rnvp_optim.Adam(m, lr=5e-4, weight_decay=5e-5)
    ref = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=5e-5)
    g, r = opt.param_groups[0], ref.param_groups[0]
    assert len(g["params"]) == len(r["params"]) == len(list(m.parameters()))
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert g[k] == r[k], k
    assert opt.state_dict()["param_groups"][0]["params"] == ref.state_dict()["param_groups"][0]["params"]
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    with pytest.raises(TypeError):
        pkg.rnvp_optim.Adam(torch.nn.Linear(2, 2))


def test_non_default_hps_construct_like_the_reference(pkg, golden_dir):
    """Every hyper-parameter branch builds (module tree, parameter order, requires_grad flags as the reference, see
    oracle/make_golden_api.py for the value checks on the GPU); on CPU tensors they refuse to run like the rest."""
    prior = torch.distributions.Normal(torch.tensor(0.), torch.tensor(1.))
    for bott, skip, wn, cbn, R in [(False, True, True, True, 1), (True, False, True, True, 1), (True, True, False, True, 1),
                                   (True, True, True, False, 1), (True, True, True, True, 0), (False, False, False, False, 0)]:
        m = pkg.RealNVP(3, 32, prior, pkg.Hyperparameters(4, R, bott, skip, wn, cbn))
        names = [n for n, _ in m.named_parameters()]
        assert any(n.endswith("conv.weight") for n in names) == (not wn)
        assert any(".core_skips." in n for n in names) == (skip and R > 0)
        assert any(".res_block.6." in n for n in names) == (bott and R > 0)
        assert any(n.startswith("s1_ckbd.0.block.1.block.") for n in names) == (R == 0)
        with pytest.raises(RuntimeError, match="no CPU path"):
            m(torch.zeros(1, 3, 32, 32))
        with pytest.raises(RuntimeError, match="no CPU path"):
            m.s1_ckbd[0](torch.zeros(1, 3, 32, 32))


def test_order_matrix_matches_index_map(pkg):
    m = _model(pkg)
    w = m.order_matrix(3)
    t = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).reshape(2, 3, 8, 8)
    full = torch.nn.functional.conv2d(t, w, stride=2)
    on, off = O.factor_out(t)
    assert torch.equal(full, torch.cat((on, off), 1))
