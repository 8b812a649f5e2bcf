"""GPU debug helper: BN-prologue conv in train mode, dump the saved coefficients and error maps."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
pkg = importlib.import_module("dl-normalizing-flows_b200")
import test_gpu_ops as T
lib, check, ptr = pkg.rnvp_cabi.lib, pkg.rnvp_cabi.check, pkg.rnvp_cabi.ptr
DEV = "cuda"
for shape in [(2, 64, 32, 32, 1), (1, 32, 64, 64, 1), (2, 16, 128, 128, 1)]:
    for mode in (1, 0):
        B, S, cin, cout, k = shape
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, cin, S, S, generator=g) * 1.5 + 0.3
        v = torch.randn(cout, cin, k, k, generator=g); gg = torch.rand(cout, 1, 1, 1, generator=g) + 0.5
        w = v * (gg / torch.linalg.vector_norm(v, dim=(1, 2, 3), keepdim=True))
        gamma, beta = torch.rand(cin, generator=g) + 0.5, torch.rand(cin, generator=g) * 0.6
        rm0, rv0 = torch.randn(cin, generator=g) * 0.1 + 0.3, torch.rand(cin, generator=g) + 1.5
        mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
        if mode == 0: mean, var = rm0, rv0
        rstd = 1 / torch.sqrt(var + 1e-5)
        h = F.relu((x - mean.view(1, -1, 1, 1)) * (gamma * rstd).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1))
        y_ref = F.conv2d(h, w, None, padding=k // 2)
        kpad, npad, ldy = T._pad(cin, 32), T._pad(cout, 16), T._pad(cout, 32)
        wf, _ = T._wn_operands(pkg, v, gg)
        xn = T._nhwc(x, kpad); y = torch.zeros(B, S, S, ldy, device=DEV)
        P = B * S * S
        sums = torch.cat((x.double().sum((0, 2, 3)), (x.double() ** 2).sum((0, 2, 3)))).to(DEV)
        rm, rv, save = rm0.to(DEV), rv0.to(DEV), torch.zeros(4 * cin, device=DEV)
        gd, bd = gamma.to(DEV), beta.to(DEV)
        torch.cuda.synchronize()
        check(lib.rnvp_conv_forward_bn(ptr(xn), ptr(wf), None, None, ptr(y), None, B, S, kpad, cout, npad, k, ldy, mode, cin,
                                       ptr(sums), float(P), ptr(gd), ptr(bd), ptr(rm), ptr(rv), ptr(save), 0, T._stream()))
        torch.cuda.synchronize()
        got = y[..., :cout].permute(0, 3, 1, 2).cpu()
        err = (got - y_ref).abs()
        sv = save.cpu().view(4, cin)
        print(shape, "mode", mode, "rel", float(err.max() / y_ref.abs().max()), "err by image", [float(e.max()) for e in err],
              "| err by row-block", [float(err[:, :, i:i + max(1, S // 4)].max()) for i in range(0, S, max(1, S // 4))])
        if mode == 1:
            print("   save mean err", float((sv[0] - mean).abs().max()), "rstd err", float((sv[1] - rstd).abs().max()),
                  "scale err", float((sv[2] - gamma * rstd).abs().max()), "shift err", float((sv[3] - (beta - mean * gamma * rstd)).abs().max()))
