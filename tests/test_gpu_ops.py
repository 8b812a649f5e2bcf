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
    check(lib.rnvp_conv_forward(ptr(xn), ptr(wf), ptr(bias.to(DEV)), ptr(y) if with_res else None, ptr(y), ptr(stats),
                                B, S, kpad, cout, npad, k, ldy, math, _stream()))
    got = y[..., :cout].permute(0, 3, 1, 2).cpu()
    tol = 1e-5 if math == 0 else 3e-3
    assert rel(got, y_ref) < tol, (rel(got, y_ref), B, S, cin, cout, k)
    s_ref = torch.cat((y_ref.double().sum((0, 2, 3)), (y_ref.double() ** 2).sum((0, 2, 3))))
    assert rel(stats, s_ref) < (1e-5 if math == 0 else 3e-3)
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
    assert rel(got_dw, dw_ref) < (2e-5 if math == 0 else 3e-3), rel(got_dw, dw_ref)
    assert rel(db, dy.sum((0, 2, 3))) < (2e-5 if math == 0 else 2e-3)   # tf32: rides on an MMA
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


HALO_SHAPES = [(2, 64, 32, 32, 3), (2, 32, 64, 64, 3), (2, 16, 128, 128, 3), (3, 16, 25, 128, 3), (2, 64, 7, 32, 3),
               (1, 32, 13, 64, 3)]


def test_conv_tf32_halo():
    """The opt-in 3x3 halo path (RNVP_HALO=1, read once per process): one haloed activation tile per 32-channel
    chunk, the nine taps through row-shifted shared-memory descriptors; streamed and resident weight tiles."""
    import os
    import subprocess
    import sys
    code = ("import sys, importlib; sys.path[:0] = [%r, %r]; import conftest; "
            "pkg = importlib.import_module('dl-normalizing-flows_b200'); import test_gpu_ops as T; "
            "[T._conv_case(pkg, *s, 1) for s in T.HALO_SHAPES]; print('halo ok')"
            % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    env = dict(os.environ, RNVP_HALO="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "halo ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
