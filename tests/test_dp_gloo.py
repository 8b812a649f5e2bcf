"""world_size-2 gloo tests (CPU) of the data-parallel host logic, plus the synchronised-BN math.

The CUDA kernels cannot run here; what can be checked without a GPU is (i) the rendezvous helpers of
rnvp_dp (unique-id exchange, state broadcast, batch sharding, bucket planning) over a real
torch.distributed group, and (ii) that the statistic-merge scheme the runtime uses -- all-reduce of
per-channel (sum, sum of squares) forward, (sum g, sum g*xhat) backward, gradient average -- reproduces
the single-process result; (ii) runs the oracle's arithmetic on two ranks."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "dl-normalizing-flows_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return [ret[r] for r in range(world)]


def _host_logic(rank, world):
    import rnvp_dp
    # unique id: rank 0's 128 bytes reach everybody
    uid = rnvp_dp.exchange_unique_id(lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    # state broadcast is in place and makes ranks identical
    torch.manual_seed(rank)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    ptrs = [p.data_ptr() for p in m.parameters()]
    rnvp_dp.broadcast_state(m)
    assert ptrs == [p.data_ptr() for p in m.parameters()]
    flat = torch.cat([t.flatten().float() for t in list(m.parameters()) + list(m.buffers())])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # uneven sharding (communication-free work only) covers the batch exactly once
    sl = [rnvp_dp.shard_batch(11, r, world, drop_last=False) for r in range(world)]
    assert sl[0][0] == 0 and sl[-1][1] == 11 and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    # the training default: equal shards, the n % world trailing samples are dropped
    eq = [rnvp_dp.shard_batch(11, r, world) for r in range(world)]
    assert len({b - a for a, b in eq}) == 1 and eq[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(eq, eq[1:]))
    return rnvp_dp.shard_batch(11, rank, world, drop_last=False), rnvp_dp.shard_batch(11, rank, world)


def test_dp_host_logic_gloo():
    out = _run(_host_logic)
    assert out == [((0, 6), (0, 5)), ((6, 11), (5, 10))]


def test_bucket_planning():
    sys.path.insert(0, os.path.join(ROOT, "dl-normalizing-flows_b200"))
    import rnvp_dp
    # cfg A-like: tiny early couplings, big late ones (forward order)
    sizes = [5, 5, 5, 20, 20, 20, 300, 300, 2000, 2000]
    ranges, o = [], 0
    for s in sizes:
        ranges.append((o, o + s))
        o += s
    b = rnvp_dp.plan_buckets(ranges, 500)
    assert b[0] == (ranges[9][0], ranges[9][1], 9) and b[1] == (ranges[8][0], ranges[8][1], 8)
    assert b[2] == (ranges[6][0], ranges[7][1], 6)
    assert b[-1][0] == 0                                   # the tail is flushed at coupling 0
    covered = sorted((x[0], x[1]) for x in b)
    assert covered[0][0] == 0 and covered[-1][1] == o
    assert all(a[1] == c[0] for a, c in zip(covered, covered[1:]))
    assert rnvp_dp.plan_buckets([(0, 10)], 1 << 20) == [(0, 10, 0)]


def _syncbn_math(rank, world):
    """One coupling on two ranks with merged statistics == the same coupling on the full batch."""
    import realnvp_oracle as O
    import torch.nn.functional as F

    C, S, D, R = 6, 4, 8, 1
    shapes = O.coupling_state_shapes("c", "ckbd", C, D, R)
    st0 = O.random_state_from_shapes(shapes, seed=3)
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(8, C, S, S, generator=g)
    w_all = torch.randn(8, generator=g)

    # --- reference: single process, whole batch ---------------------------------------------------
    st = {k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st0.items()}
    ora = O.RealNVPOracle(st, C, S, D, R, 1)
    ora.update_running = False
    y, J = ora.coupling("c", x_all, kind="ckbd", cfg=1)
    loss = ((y ** 2).sum((1, 2, 3)) * 0.1 + J.sum((1, 2, 3))) @ w_all / 8
    loss.backward()
    ref = {k: v.grad.clone() for k, v in st.items() if v.grad is not None}

    # --- two ranks: batch statistics merged by all-reduce of (sum, sumsq), as the runtime does -----
    b0, b1 = rank * 4, rank * 4 + 4
    orig_bn = F.batch_norm

    class SyncBN(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, eps):
            n = torch.tensor(float(x.numel() // x.shape[1]))
            s = torch.stack((x.double().sum((0, 2, 3)), (x.double() ** 2).sum((0, 2, 3))))
            dist.all_reduce(s)
            dist.all_reduce(n)
            mean = (s[0] / n).float()
            var = (s[1] / n - (s[0] / n) ** 2).clamp_min(0).float()
            rstd = 1.0 / torch.sqrt(var + eps)
            xh = (x - mean[None, :, None, None]) * rstd[None, :, None, None]
            ctx.save_for_backward(xh, w, rstd)
            ctx.n = n
            out = xh
            if w is not None:
                out = out * w[None, :, None, None] + b[None, :, None, None]
            return out

        @staticmethod
        def backward(ctx, gy):
            xh, w, rstd = ctx.saved_tensors
            gh = gy * w[None, :, None, None] if w is not None else gy
            s = torch.stack((gh.double().sum((0, 2, 3)), (gh.double() * xh.double()).sum((0, 2, 3))))
            dist.all_reduce(s)                                  # (sum g, sum g*xhat) over ALL ranks
            m1, m2 = (s[0] / ctx.n).float(), (s[1] / ctx.n).float()
            gx = rstd[None, :, None, None] * (gh - m1[None, :, None, None] - xh * m2[None, :, None, None])
            gw = (gy * xh).sum((0, 2, 3)) if w is not None else None
            gb = gy.sum((0, 2, 3)) if w is not None else None
            return gx, gw, gb, None

    def fake_bn(x, rm, rv, w, b, training, momentum, eps):
        assert training
        return SyncBN.apply(x, w, b, eps)

    F.batch_norm = fake_bn
    try:
        st2 = {k: v.clone().requires_grad_(O.is_trainable(k) and v.is_floating_point()) for k, v in st0.items()}
        ora2 = O.RealNVPOracle(st2, C, S, D, R, 1)
        ora2.update_running = False
        x = x_all[b0:b1]
        # the coupling's own batch_stat for the log-det term must be global as well
        orig_mean = torch.mean

        def global_mean(t, dim=None, keepdim=False):
            if dim == (0, 2, 3):
                s = t.sum(dim=dim, keepdim=keepdim)
                s = _AllReduceSum.apply(s)
                return s / (t.numel() // t.shape[1] * world)
            return orig_mean(t, dim=dim, keepdim=keepdim) if dim is not None else orig_mean(t)

        class _AllReduceSum(torch.autograd.Function):
            @staticmethod
            def forward(ctx, s):
                s = s.clone()
                dist.all_reduce(s)
                return s

            @staticmethod
            def backward(ctx, gs):
                gs = gs.clone()
                dist.all_reduce(gs)
                return gs

        torch.mean = global_mean
        try:
            y2, J2 = ora2.coupling("c", x, kind="ckbd", cfg=1)
        finally:
            torch.mean = orig_mean
        loss2 = ((y2 ** 2).sum((1, 2, 3)) * 0.1 + J2.sum((1, 2, 3))) @ w_all[b0:b1] / 4    # local mean
        loss2.backward()
    finally:
        F.batch_norm = orig_bn
    worst = 0.0
    gmax = max(float(v.abs().max()) for v in ref.values())
    for k, gr in ref.items():
        gl = st2[k].grad.clone()
        dist.all_reduce(gl)
        gl /= world                                           # the runtime's ncclAvg
        worst = max(worst, float((gl - gr).abs().max()) / max(float(gr.abs().max()), 1e-3 * gmax))
    assert torch.allclose(y2, y[b0:b1].detach(), rtol=1e-4, atol=1e-5)
    return worst


def test_syncbn_merge_reproduces_single_process():
    out = _run(_syncbn_math)
    assert max(out) < 5e-3, out
