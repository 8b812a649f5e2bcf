"""Host-side cost of one training step: wall time of the enqueue (no synchronisation) against the device time.
GPU box only.  Usage: [B=256] python tools/host_time.py"""
import importlib, os, sys, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
pkg = importlib.import_module("dl-normalizing-flows_b200")
B = int(os.environ.get("B", "256"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
prior = torch.distributions.Normal(torch.tensor(0., device=dev), torch.tensor(1., device=dev), validate_args=False)
model = pkg.RealNVP(3, 64, prior, pkg.Hyperparameters(32, 4, True, True, True, True)).to(dev)
model.set_math("tf32")
opt = pkg.rnvp_optim.Adam(model, lr=5e-4, weight_decay=5e-5)
x_u8 = torch.randint(0, 256, (B, 3, 64, 64), dtype=torch.uint8, device=dev)
model.train()


def step():
    t = [time.perf_counter()]
    opt.zero_grad(set_to_none=False)
    x, logdet = pkg.logit_transform(x_u8)
    ll, ws = model(x)
    t.append(time.perf_counter())
    loss = -(ll + logdet).mean() + 5e-5 * ws
    loss.backward()
    t.append(time.perf_counter())
    opt.step()
    t.append(time.perf_counter())
    return t


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acc = [0.0, 0.0, 0.0]
e0.record()
w0 = time.perf_counter()
for _ in range(n):
    t = step()
    for i in range(3):
        acc[i] += t[i + 1] - t[i]
w1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
print(f"B={B}: device {e0.elapsed_time(e1) / n:.2f} ms/step; host enqueue {1e3 * (w1 - w0) / n:.2f} ms/step "
      f"(forward {1e3 * acc[0] / n:.2f}, backward {1e3 * acc[1] / n:.2f}, optimizer {1e3 * acc[2] / n:.2f}); "
      f"launches/step {pkg.rnvp_cabi.lib.rnvp_launch_count() // (n + 3)}")
# with a synchronisation before every step the host cannot run ahead: enqueue time is then pure host cost
acc = [0.0, 0.0, 0.0]
for _ in range(n):
    torch.cuda.synchronize()
    t = step()
    for i in range(3):
        acc[i] += t[i + 1] - t[i]
torch.cuda.synchronize()
print(f"      with a sync before each step: forward {1e3 * acc[0] / n:.2f}, backward {1e3 * acc[1] / n:.2f}, "
      f"optimizer {1e3 * acc[2] / n:.2f} ms of host time")
