"""-m gpu: building-block kernels through the C-ABI against the oracle / golden fixtures."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

import realnvp_oracle as O
from _util import rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_device_is_b200(pkg):
    pkg.rnvp_cabi.check(pkg.rnvp_cabi.lib.rnvp_device_ok())


def test_logit_golden(pkg, golden_dir):
    lg = torch.load(os.path.join(golden_dir, "misc.pt"))["logit"]
    y, ld = pkg.logit_transform(lg["x"].to(DEV), noise=lg["noise"].to(DEV))
    assert rel(y, lg["y"]) < 1e-6
    assert rel(ld, lg["logdet"]) < 1e-6
    inv, zero = pkg.logit_transform(lg["y"].to(DEV), reverse=True)
    assert zero == 0 and rel(inv, lg["inv"]) < 1e-6
    # uint8 entry point == float entry point on value/255
    x8 = (lg["x"] * 255).round().to(torch.uint8)
    y8, ld8 = pkg.logit_transform(x8.to(DEV), noise=lg["noise"].to(DEV))
    yo, ldo = O.logit_forward(x8.float() / 255.0, lg["noise"])
    assert rel(y8, yo) < 1e-6 and rel(ld8, ldo) < 1e-6


def test_logit_inkernel_noise_and_edges(pkg):
    x = O.synthetic_images(8, 3, 64, seed=2).to(DEV)
    torch.manual_seed(5)
    y1, ld1 = pkg.logit_transform(x)
    torch.manual_seed(5)
    y2, _ = pkg.logit_transform(x)
    torch.manual_seed(6)
    y3, _ = pkg.logit_transform(x)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    # the implied noise is U[0,1): invert the transform and check the range and the mean
    xin = torch.sigmoid(y1.double())
    u = ((xin * 2 - 1) / 0.9 + 1) / 2 * 256 - x.double() * 255
    assert float(u.min()) > -1e-3 and float(u.max()) < 1 + 1e-3 and abs(float(u.mean()) - 0.5) < 0.01
    # per-sample log-det equals the oracle's on the recovered noise
    _, ldo = O.logit_forward(x.cpu(), u.clamp(0, 1).float().cpu())
    assert rel(ld1, ldo) < 1e-4
    # empty batch
    y0, ld0 = pkg.logit_transform(torch.zeros(0, 3, 8, 8, device=DEV))
    assert y0.shape == (0, 3, 8, 8) and ld0.shape == (0,)


def test_layout_golden(pkg, golden_dir):
    lay = torch.load(os.path.join(golden_dir, "misc.pt"))["layout"]
    prior = torch.distributions.Normal(torch.tensor(0.), torch.tensor(1.))
    m = pkg.RealNVP(3, 32, prior, pkg.Hyperparameters(4, 1, True, True, True, True))
    t = lay["t"].to(DEV)
    sq = m.squeeze(t)
    assert torch.equal(sq.cpu(), lay["squeeze"])
    assert torch.equal(m.undo_squeeze(sq).cpu(), lay["t"])
    on, off = m.factor_out(t, m.order_matrix_1)
    assert torch.equal(on.cpu(), lay["on"]) and torch.equal(off.cpu(), lay["off"])
    assert torch.equal(m.restore(on, off, m.order_matrix_1).cpu(), lay["t"])
    # bigger, odd channel count, full-size spatial
    g = torch.Generator().manual_seed(0)
    t2 = torch.randn(5, 7, 64, 64, generator=g)
    on2, off2 = m.factor_out(t2.to(DEV))
    oo, of = O.factor_out(t2)
    assert torch.equal(on2.cpu(), oo) and torch.equal(off2.cpu(), of)
    assert torch.equal(m.squeeze(t2.to(DEV)).cpu(), O.squeeze(t2))


def _pad(n, m):
    return (n + m - 1) // m * m


def _with_lo(w):
    """[w | w - trunc_tf32(w)]: the weight operand of the 3xTF32 tier (math = 2) through the per-op entry points
    (inside the library the weight-norm kernel writes both copies)."""
    hi = (w.view(torch.int32) & -8192).view(torch.float32)          # 0xffffe000: what kind::tf32 reads from an fp32 word
    return torch.cat((w.reshape(-1), (w - hi).reshape(-1)))


def _conv_case(pkg, B, S, cin, cout, k, math, seed=0, with_res=True):
    lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
    g = torch.Generator().manual_seed(seed)
    kpad, npad = _pad(cin, 32), _pad(cout, 16)
    kpad_b, npad_b = _pad(cout, 32), _pad(cin, 16)
    v = torch.randn(cout, cin, k, k, generator=g)
    gg = torch.rand(cout, 1, 1, 1, generator=g) + 0.5
    x = torch.randn(B, cin, S, S, generator=g)
    bias = torch.randn(cout, generator=g)
    res = torch.randn(B, cout, S, S, generator=g)
    w = v * (gg / torch.linalg.vector_norm(v, dim=(1, 2, 3), keepdim=True))
    y_ref = F.conv2d(x, w, bias, padding=k // 2) + (res if with_res else 0)
    # device operands
    warena = torch.empty(k * k * npad * kpad + k * k * npad_b * kpad_b, device=DEV)
    wf, wb = warena[: k * k * npad * kpad], warena[k * k * npad * kpad:]
    vd, gd = v.to(DEV), gg.to(DEV)
    check(lib.rnvp_weightnorm_forward(ptr(vd), ptr(gd), ptr(wf), ptr(wb), cout, cin, k, _stream()))
    wf_ref = torch.zeros(k * k, npad, kpad)
    wf_ref[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    assert rel(wf.view(k * k, npad, kpad), wf_ref) < 1e-6
    wb_ref = torch.zeros(k * k, npad_b, kpad_b)
    wb_ref[:, :cin, :cout] = w.flip(2, 3).permute(2, 3, 1, 0).reshape(k * k, cin, cout)
    assert rel(wb.view(k * k, npad_b, kpad_b), wb_ref) < 1e-6
    xn = torch.zeros(B, S, S, kpad)
    xn[..., :cin] = x.permute(0, 2, 3, 1)
    xn = xn.to(DEV)
    ldy = _pad(cout, 32)
    resn = torch.zeros(B, S, S, ldy)
    resn[..., :cout] = res.permute(0, 2, 3, 1)
    y = resn.to(DEV).clone()
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    bias_d = bias.to(DEV)
    exact = math in (0, 2)                       # CUDA-core fp32 and 3xTF32 on the tensor cores: the 1e-5 tier
    if math == 2:
        wf, wb = _with_lo(wf), _with_lo(wb)
    check(lib.rnvp_conv_forward(ptr(xn), ptr(wf), ptr(bias_d), ptr(y) if with_res else None, ptr(y), ptr(stats),
                                B, S, kpad, cout, npad, k, ldy, math, _stream()))
    got = y[..., :cout].permute(0, 3, 1, 2).cpu()
    # 3xTF32 (measured 3e-7 ... 4.1e-6 of max, the largest at K = 4608: the tensor core truncates its fp32 accumulator
    # after every MMA, which is why that tier spreads the products over partial accumulators) meets the fp32 tolerance
    tol = 1e-5 if exact else 3e-3
    if math == 2:
        print(f"[tf32x3] conv {B}x{S}x{S} {cin}->{cout} k{k}: {rel(got, y_ref):.2e} of max")
    assert rel(got, y_ref) < tol, (rel(got, y_ref), B, S, cin, cout, k)
    s_ref = torch.cat((y_ref.double().sum((0, 2, 3)), (y_ref.double() ** 2).sum((0, 2, 3))))
    assert rel(stats, s_ref) < (tol if exact else 3e-3)
    if with_res:
        assert float(y[..., cout:].abs().max()) == 0.0 if ldy > cout else True
    # dgrad through the same kernel with the transposed / flipped operand
    dy = torch.randn(B, cout, S, S, generator=g)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, w, dy, padding=k // 2)
    dyn = torch.zeros(B, S, S, kpad_b)
    dyn[..., :cout] = dy.permute(0, 2, 3, 1)
    dyn = dyn.to(DEV)
    ldx = _pad(cin, 32)
    dx = torch.zeros(B, S, S, ldx, device=DEV)
    check(lib.rnvp_conv_forward(ptr(dyn), ptr(wb), None, None, ptr(dx), None, B, S, kpad_b, cin, npad_b, k, ldx, math,
                                _stream()))
    assert rel(dx[..., :cin].permute(0, 3, 1, 2), dx_ref) < tol
    # wgrad + bias grad, then weight-norm backward
    dw_ref = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=k // 2)
    dwf = torch.zeros(k * k, npad, kpad, device=DEV)
    db = torch.zeros(cout, device=DEV)
    check(lib.rnvp_conv_wgrad(ptr(xn), ptr(dyn), ptr(dwf), ptr(db), B, S, kpad, cout, npad, k, kpad_b, math, _stream()))
    got_dw = dwf[:, :cout, :cin].reshape(k, k, cout, cin).permute(2, 3, 0, 1).cpu()
    if math == 2:
        print(f"[tf32x3] dgrad {rel(dx[..., :cin].permute(0, 3, 1, 2), dx_ref):.2e}, wgrad {rel(got_dw, dw_ref):.2e} of max")
    assert rel(got_dw, dw_ref) < (2e-5 if exact else 3e-3), rel(got_dw, dw_ref)
    assert rel(db, dy.sum((0, 2, 3))) < (2e-5 if exact else 2e-3)
    vr, gr = v.clone().requires_grad_(True), gg.clone().requires_grad_(True)
    wr = vr * (gr / torch.linalg.vector_norm(vr, dim=(1, 2, 3), keepdim=True))
    (wr * dw_ref).sum().backward()
    dv, dg = torch.zeros_like(vd), torch.zeros_like(gd)
    dwf_exact = torch.zeros(k * k, npad, kpad)
    dwf_exact[:, :cout, :cin] = dw_ref.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    check(lib.rnvp_weightnorm_backward(ptr(vd), ptr(gd), ptr(dwf_exact.to(DEV)), ptr(dv), ptr(dg), cout, cin, k, _stream()))
    assert rel(dv, vr.grad) < 1e-5 and rel(dg, gr.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 8, 7, 32, 3), (3, 4, 32, 32, 1), (2, 16, 64, 6, 1), (1, 4, 97, 64, 3),
                                   (5, 2, 12, 8, 3), (2, 64, 32, 32, 3)])
def test_conv_fp32(pkg, shape):
    B, S, cin, cout, k = shape
    _conv_case(pkg, B, S, cin, cout, k, 0)
    _conv_case(pkg, B, S, cin, cout, k, 0, seed=1, with_res=False)


@pytest.mark.parametrize("shape", [(2, 8, 7, 32, 3), (3, 4, 32, 32, 1), (2, 16, 64, 6, 1), (1, 4, 97, 64, 3),
                                   (5, 2, 12, 8, 3), (2, 64, 32, 32, 3), (4, 8, 256, 512, 1), (2, 32, 64, 64, 3),
                                   (3, 8, 160, 96, 3), (40, 4, 512, 512, 1),
                                   (2, 16, 128, 128, 3), (3, 16, 25, 128, 3), (2, 64, 7, 32, 3), (1, 32, 13, 64, 3)])
def test_conv_tf32(pkg, shape):
    """tcgen05 kind::tf32 implicit GEMM (TMA-fed, TMEM accumulator) vs the fp32 torch reference."""
    B, S, cin, cout, k = shape
    _conv_case(pkg, B, S, cin, cout, k, 1)
    _conv_case(pkg, B, S, cin, cout, k, 1, seed=1, with_res=False)


@pytest.mark.parametrize("shape", [(2, 8, 7, 32, 3), (3, 4, 32, 32, 1), (2, 16, 64, 6, 1), (1, 4, 97, 64, 3),
                                   (5, 2, 12, 8, 3), (2, 64, 32, 32, 3), (4, 8, 256, 512, 1), (2, 32, 64, 64, 3),
                                   (3, 8, 160, 96, 3), (40, 4, 512, 512, 1), (21, 4, 512, 512, 3),
                                   (2, 16, 128, 128, 3), (3, 16, 25, 128, 3), (2, 64, 7, 32, 3), (1, 32, 13, 64, 3)])
def test_conv_tf32x3(pkg, shape):
    """The fp32-accurate tensor-core tier: the tcgen05 kernels with 3xTF32 split operands (hi*hi + lo*hi + hi*lo),
    conv / dgrad / wgrad / bias gradient against the fp32 torch reference at the fp32 tier's tolerances."""
    B, S, cin, cout, k = shape
    _conv_case(pkg, B, S, cin, cout, k, 2)
    _conv_case(pkg, B, S, cin, cout, k, 2, seed=1, with_res=False)


def _nhwc(t, ld):
    """(B,C,S,S) -> zero-padded NHWC [B,S,S,ld] on the device."""
    B, Cc, S, _ = t.shape
    out = torch.zeros(B, S, S, ld)
    out[..., :Cc] = t.permute(0, 2, 3, 1)
    return out.to(DEV)


def _wn_operands(pkg, v, gg):
    lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
    cout, cin, k, _ = v.shape
    kpad, npad, kpad_b, npad_b = _pad(cin, 32), _pad(cout, 16), _pad(cout, 32), _pad(cin, 16)
    arena = torch.empty(k * k * npad * kpad + k * k * npad_b * kpad_b, device=DEV)
    wf, wb = arena[: k * k * npad * kpad], arena[k * k * npad * kpad:]
    vd, gd = v.to(DEV), gg.to(DEV)       # named: a temporary's block could be handed to the next allocation (and be
    check(lib.rnvp_weightnorm_forward(ptr(vd), ptr(gd), ptr(wf), ptr(wb), cout, cin, k, _stream()))   # overwritten by
    torch.cuda.current_stream().synchronize()                                     # its H2D copy) before the kernel ran
    return wf, wb


BN_SHAPES = [(2, 64, 32, 32, 1), (2, 64, 32, 32, 3), (3, 32, 64, 64, 3), (2, 16, 128, 128, 1), (2, 16, 128, 24, 1),
             (5, 8, 256, 256, 3), (37, 4, 512, 512, 1), (9, 4, 512, 96, 1), (3, 8, 12, 32, 3), (1, 32, 64, 64, 1)]


@pytest.mark.parametrize("shape", BN_SHAPES)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_conv_bn_prologue(pkg, shape, mode):
    """BatchNorm2d -> ReLU -> conv in ONE tcgen05 kernel (the BN is applied to the staged operand tiles) against
    torch's three ops (modules_realnvp.py:83-97).  beta is shifted up so that relu(bn(0)) != 0: zero padding
    must be applied AFTER the activation.  Also pins the saved coefficients, the running-statistic update and the
    wgrad that rebuilds relu(bn(x)) from the raw x boxes."""
    lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
    B, S, cin, cout, k = shape
    g = torch.Generator().manual_seed(B * 1000 + S + cin + cout + k)
    x = torch.randn(B, cin, S, S, generator=g) * 1.5 + 0.3
    v = torch.randn(cout, cin, k, k, generator=g)
    gg = torch.rand(cout, 1, 1, 1, generator=g) + 0.5
    w = v * (gg / torch.linalg.vector_norm(v, dim=(1, 2, 3), keepdim=True))
    gamma, beta = torch.rand(cin, generator=g) + 0.5, torch.rand(cin, generator=g) * 0.6
    rm0, rv0 = torch.randn(cin, generator=g) * 0.1 + 0.3, torch.rand(cin, generator=g) + 1.5
    bias, res = torch.randn(cout, generator=g), torch.randn(B, cout, S, S, generator=g)
    bn = torch.nn.BatchNorm2d(cin)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm0); bn.running_var.copy_(rv0)
    bn.train(mode == "train")
    h_ref = F.relu(bn(x)).detach()
    y_ref = F.conv2d(h_ref, w, bias, padding=k // 2) + res
    kpad, npad, ldy = _pad(cin, 32), _pad(cout, 16), _pad(cout, 32)
    wf, _wb = _wn_operands(pkg, v, gg)
    xn = _nhwc(x, kpad)
    y = _nhwc(res, ldy)
    P = B * S * S
    sums = torch.cat((x.double().sum((0, 2, 3)), (x.double() ** 2).sum((0, 2, 3)))).to(DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    rm, rv, save = rm0.to(DEV), rv0.to(DEV), torch.zeros(4 * cin, device=DEV)
    bias_d, gamma_d, beta_d = bias.to(DEV), gamma.to(DEV), beta.to(DEV)      # named: the pointers must stay valid
    check(lib.rnvp_conv_forward_bn(ptr(xn), ptr(wf), ptr(bias_d), ptr(y), ptr(y), ptr(stats), B, S, kpad, cout, npad,
                                   k, ldy, 1 if mode == "train" else 0, cin, ptr(sums), float(P), ptr(gamma_d),
                                   ptr(beta_d), ptr(rm), ptr(rv), ptr(save), 0, _stream()))
    got = y[..., :cout].permute(0, 3, 1, 2).cpu()
    assert rel(got, y_ref) < 3e-3, (rel(got, y_ref), shape, mode)
    s_ref = torch.cat((y_ref.double().sum((0, 2, 3)), (y_ref.double() ** 2).sum((0, 2, 3))))
    assert rel(stats, s_ref) < 3e-3
    if mode == "eval":
        assert torch.equal(rm.cpu(), rm0) and torch.equal(rv.cpu(), rv0)
        return
    assert torch.allclose(rm.cpu(), bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.cpu(), bn.running_var, rtol=1e-5, atol=1e-6)
    mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
    rstd = 1 / torch.sqrt(var + 1e-5)
    sv = save.cpu().view(4, cin)
    assert rel(sv[0], mean) < 1e-5 and rel(sv[1], rstd) < 1e-5
    assert rel(sv[2], gamma * rstd) < 1e-5 and rel(sv[3], beta - mean * gamma * rstd) < 1e-5
    # the same tile transform with the coefficients read back (mode 2) reproduces the result bit for bit
    y2 = _nhwc(res, ldy)
    check(lib.rnvp_conv_forward_bn(ptr(xn), ptr(wf), ptr(bias_d), ptr(y2), ptr(y2), None, B, S, kpad, cout, npad,
                                   k, ldy, 2, cin, None, float(P), None, None, None, None, ptr(save), 0, _stream()))
    assert torch.equal(y2, y)
    # weight gradient from the RAW x boxes
    dy = torch.randn(B, cout, S, S, generator=g)
    dw_ref = torch.nn.grad.conv2d_weight(h_ref, w.shape, dy, padding=k // 2)
    dyn = _nhwc(dy, _pad(cout, 32))
    dwf = torch.zeros(k * k, npad, kpad, device=DEV)
    db = torch.zeros(cout, device=DEV)
    check(lib.rnvp_conv_wgrad_bn(ptr(xn), ptr(dyn), ptr(dwf), ptr(db), B, S, kpad, cout, npad, k, _pad(cout, 32),
                                 ptr(save), cin, _stream()))
    got_dw = dwf[:, :cout, :cin].reshape(k, k, cout, cin).permute(2, 3, 0, 1).cpu()
    assert rel(got_dw, dw_ref) < 3e-3, (rel(got_dw, dw_ref), shape)
    assert rel(db, dy.sum((0, 2, 3))) < 2e-3
    assert float(dwf[:, cout:, :].abs().max() if npad > cout else 0) == 0.0


@pytest.mark.parametrize("shape", [(2, 64, 32, 32, 1), (2, 32, 64, 64, 3), (3, 16, 128, 128, 3), (5, 8, 256, 256, 1),
                                   (21, 4, 512, 512, 3), (2, 16, 24, 128, 1), (3, 8, 96, 256, 1)])
def test_dgrad_bn_relu_fused(pkg, shape):
    """The backward unit of x -> BN(train) -> ReLU -> conv: dgrad with the ReLU mask and the two BN-backward
    reductions in its epilogue (rnvp_conv_dgrad_bn), then rnvp_bn_backward_apply -- against torch autograd through
    the three reference ops (modules_realnvp.py:83-97), incl. dgamma / dbeta and the residual-gradient add."""
    lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
    B, S, cout, cin, k = shape                 # conv: cin -> cout; the BN under test normalises its cin-channel INPUT
    g = torch.Generator().manual_seed(7 + B + S + cin + cout + k)
    x = (torch.randn(B, cin, S, S, generator=g) * 1.3 + 0.2).requires_grad_(True)
    v = torch.randn(cout, cin, k, k, generator=g)
    gg = torch.rand(cout, 1, 1, 1, generator=g) + 0.5
    w = v * (gg / torch.linalg.vector_norm(v, dim=(1, 2, 3), keepdim=True))
    gamma = (torch.rand(cin, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.rand(cin, generator=g) * 0.4 - 0.2).requires_grad_(True)
    dy = torch.randn(B, cout, S, S, generator=g)
    add = torch.randn(B, cin, S, S, generator=g)
    h = F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5))
    (F.conv2d(h, w, None, padding=k // 2) * dy).sum().backward()
    P = B * S * S
    xd = x.detach()
    mean, var = xd.mean((0, 2, 3)), xd.var((0, 2, 3), unbiased=False)
    rstd = 1 / torch.sqrt(var + 1e-5)
    gd, bd = gamma.detach(), beta.detach()
    save = torch.cat((mean, rstd, gd * rstd, bd - mean * gd * rstd)).to(DEV)
    _wf, wb = _wn_operands(pkg, v, gg)
    ld = _pad(cin, 32)
    xn, dyn = _nhwc(xd, ld), _nhwc(dy, _pad(cout, 32))
    gm = torch.zeros(B, S, S, ld, device=DEV)
    sums2 = torch.zeros(2 * cin, dtype=torch.float64, device=DEV)
    check(lib.rnvp_conv_dgrad_bn(ptr(dyn), ptr(wb), ptr(xn), ptr(save), ptr(gm), ptr(sums2), B, S, _pad(cout, 32), cin,
                                 _pad(cin, 16), k, ld, _stream()))
    # The kernel takes the ReLU mask from fma(x, scale, shift) > 0, torch from its own BN formula: an activation within
    # ~1e-6 of the threshold may be masked differently (an O(1) difference on that one element).  Such elements are
    # excluded from the pointwise comparison and bounded in number.
    pre = (xd - mean.view(1, -1, 1, 1)) * (gd * rstd).view(1, -1, 1, 1) + bd.view(1, -1, 1, 1)
    sure = pre.abs() > 1e-5
    assert int((~sure).sum()) <= 8
    dh_ref = torch.nn.grad.conv2d_input(xd.shape, w, dy, padding=k // 2) * (h.detach() > 0)
    gm_c = gm[..., :cin].permute(0, 3, 1, 2).cpu()
    assert rel(gm_c * sure, dh_ref * sure) < 3e-3
    s_ref = torch.cat((dh_ref.double().sum((0, 2, 3)), (dh_ref.double() * xd.double()).sum((0, 2, 3))))
    scale_s = float(dh_ref.double().abs().sum((0, 2, 3)).max())          # sums of signed values: compare on the L1 scale
    assert float((sums2.cpu() - s_ref).abs().max()) < 3e-3 * scale_s
    dx = torch.zeros_like(gm)
    dgam, dbet = torch.zeros(cin, device=DEV), torch.zeros(cin, device=DEV)
    addn = _nhwc(add, ld)
    gd_d = gd.to(DEV)
    check(lib.rnvp_bn_backward_apply(ptr(gm), ptr(xn), ptr(dx), ptr(addn), P, cin, ld, ptr(save), ptr(sums2), float(P),
                                     ptr(gd_d), ptr(dgam), ptr(dbet), 1, 0, _stream()))
    got = dx[..., :cin].permute(0, 3, 1, 2).cpu() - add
    assert rel(got * sure, x.grad * sure) < 5e-3, (rel(got * sure, x.grad * sure), shape)
    # per-channel sums of TF32-noisy values with cancellation (K up to 4608 products per element)
    assert rel(dgam, gamma.grad) < 1e-2 and rel(dbet, beta.grad) < 1e-2
    if ld > cin:
        assert float(dx[..., cin:].abs().max()) == 0.0
    # the rounded variant only differs by the TF32 rounding of its output
    dxr = torch.zeros_like(gm)
    dgam.zero_(); dbet.zero_()
    check(lib.rnvp_bn_backward_apply(ptr(gm), ptr(xn), ptr(dxr), ptr(addn), P, cin, ld, ptr(save), ptr(sums2), float(P),
                                     ptr(gd_d), ptr(dgam), ptr(dbet), 1, 1, _stream()))
    assert torch.equal(dxr.cpu(), O.round_tf32(dx.cpu()))
