"""Data-parallel training for the drop-in RealNVP (one process per GPU, torchrun / torch.distributed).

The reference is single-device (train.py:103-106); this is the north star's multi-GPU row:

* parameters and buffers are broadcast from rank 0 at wrap time;
* the gradient all-reduce is done by the C-ABI runtime itself over its own NCCL communicator: the
  flat gradient buffer is averaged in buckets of whole couplings on a side stream while the backward
  pass of the earlier couplings is still running (``rnvp_dp_set_grad_layout``);
* every batch norm of the stack becomes a synchronised batch norm: the per-channel double sums are
  all-reduced in-stream between the producer and the consumer kernel, so ``log_prob`` and all
  gradients equal those of a single process running the concatenated batch;
* sampling (``sample`` / ``g`` in eval mode) needs no communication: each rank draws its own z.

``torch.distributed`` (any backend) is only used for the rendezvous: shipping the 128-byte NCCL id
and the initial parameter broadcast.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


def plan_buckets(ranges: Sequence[Tuple[int, int]], min_elems: int) -> List[Tuple[int, int, int]]:
    """Group per-coupling gradient ranges (forward order) into all-reduce buckets the way the runtime
    does while walking the couplings last -> first: a bucket closes once it holds >= min_elems
    elements or the first coupling is reached.  Returns (begin, end, first_coupling) in launch order."""
    out = []
    pending_end = ranges[-1][1] if ranges else 0
    for ci in range(len(ranges) - 1, -1, -1):
        begin = ranges[ci][0]
        if pending_end - begin >= min_elems or ci == 0:
            if pending_end > begin:
                out.append((begin, pending_end, ci))
            pending_end = begin
    return out


def exchange_unique_id(make_id, group=None, device="cpu") -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id; everybody receives it over ``group``."""
    rank = dist.get_rank(group)
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def broadcast_state(module: nn.Module, group=None) -> None:
    """All ranks start from rank 0's parameters and buffers (in place: pointers stay bound)."""
    src = dist.get_global_rank(group, 0) if group is not None else 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def shard_batch(n: int, rank: int, world: int, drop_last: bool = True) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a global batch of n samples owned by ``rank``.

    Synchronised batch norm (count = local pixels x world) and the ncclAvg gradient reduction need EQUAL local
    batches, so by default the n % world trailing samples are dropped; ``drop_last=False`` hands them out unevenly
    and is only valid for communication-free work (sampling / evaluation).  Training with unequal shards is
    detected on the device at the next forward (sticky error 3, ``DataParallel.check_health``)."""
    per, rem = divmod(n, world)
    if drop_last:
        return rank * per, rank * per + per
    begin = rank * per + min(rank, rem)
    return begin, begin + per + (1 if rank < rem else 0)


class DataParallel(nn.Module):
    """Wraps a CUDA-resident ``flow_realnvp.RealNVP``; ``forward`` / ``log_prob`` / ``sample`` delegate."""

    def __init__(self, module: nn.Module, group=None, bucket_elems: int = 1 << 20):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (torchrun) before wrapping")
        from rnvp_cabi import check, lib
        self.module = module
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DataParallel needs the model on its CUDA device first; there is no CPU path")
        broadcast_state(module, group)
        eng = module.engine()
        eng.ensure_bound(dev)

        def make_id():
            raw = (C.c_ubyte * 128)()
            check(lib.rnvp_dp_unique_id(raw))
            return bytes(raw)

        uid = exchange_unique_id(make_id, group, device=dev)
        raw = (C.c_ubyte * 128).from_buffer_copy(uid)
        check(lib.rnvp_dp_init(eng.handle, raw, self.rank, self.world))
        self.stat_exchange = self._open_stat_exchange(eng, dev)
        self._install_layout(eng, bucket_elems)
        eng.dp = self
        self._bucket_elems = bucket_elems

    def _open_stat_exchange(self, eng, dev) -> str:
        """Map every rank's statistic inbox into every other rank (CUDA IPC over NVLink) so that the 840+
        batch-norm statistic reductions of a step are one-shot peer-memory kernels instead of NCCL calls.
        Falls back to NCCL (returns "nccl") when IPC is unavailable or RNVP_DP_XCHG=0; all ranks agree."""
        import os
        from rnvp_cabi import lib
        ok = 1 if (self.world > 1 and self.world <= 16 and os.environ.get("RNVP_DP_XCHG", "1") != "0") else 0
        handle = (C.c_ubyte * 64)()
        if ok and lib.rnvp_dp_xchg_alloc(eng.handle, 2 * 1024 + 8, handle) != 0:
            ok = 0
        mine = torch.tensor(list(bytes(handle)) + [ok], dtype=torch.uint8, device=dev)
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.group)
        every = [bytes(t.cpu().numpy().tobytes()) for t in every]
        if not all(e[64] for e in every):
            return "nccl"
        blob = b"".join(e[:64] for e in every)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        opened = 1 if lib.rnvp_dp_xchg_open(eng.handle, buf) == 0 else 0
        # the exchange is collective: use it only if every rank mapped every inbox
        flag = torch.tensor([opened], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            raise RuntimeError("statistic exchange: a rank could not map its peers' inboxes (CUDA IPC); "
                               "set RNVP_DP_XCHG=0 to use NCCL for the batch-norm statistics")
        return "nvlink"

    def _install_layout(self, eng, bucket_elems):
        from rnvp_cabi import check, lib
        offs = [r[0] for r in eng.cpl_grad_ranges] + [eng.cpl_grad_ranges[-1][1]]
        arr = (C.c_int64 * len(offs))(*offs)
        check(lib.rnvp_dp_set_grad_layout(eng.handle, C.c_void_p(eng._flat_grad.data_ptr()), arr, bucket_elems))
        self._layout_ptr = eng._flat_grad.data_ptr()
        self.buckets = plan_buckets(eng.cpl_grad_ranges, bucket_elems)

    def reduce_gradients(self, eng) -> None:
        """Hook called by the engine after rnvp_flow_backward: the runtime has already enqueued the
        bucketed all-reduce; only a re-bind (moved flat buffer) needs the layout refreshed."""
        if eng._flat_grad.data_ptr() != self._layout_ptr:
            self._install_layout(eng, self._bucket_elems)
            raise RuntimeError("gradient buffer moved during a step; re-run the step")
        self.check_health()

    _XCHG_ERRORS = {1: "a peer rank never arrived at a batch-norm statistic exchange",
                    2: "the ranks issued different sequences of train-mode calls",
                    3: "the ranks' local batch sizes differ (synchronised batch norm and the gradient average need "
                       "equal shards: use drop_last / a global batch divisible by the world size)"}

    def check_health(self) -> None:
        """Raise if the statistic exchange has set its sticky error word (read from mapped host memory: free)."""
        from rnvp_cabi import lib
        e = lib.rnvp_dp_xchg_errors(self.module.engine().handle)
        if e > 0:
            raise RuntimeError(f"rnvp data parallel: {self._XCHG_ERRORS.get(e, e)}; the replicas have diverged")

    def forward(self, x):
        eng = self.module.engine()
        eng.ensure_bound(x.device)
        if eng._flat_grad.data_ptr() != self._layout_ptr:
            self._install_layout(eng, self._bucket_elems)
        self.check_health()
        return self.module(x)

    def log_prob(self, x):
        return self.forward(x)[0]

    def sample(self, size):
        return self.module.sample(size)

    def g(self, z):
        return self.module.g(z)

    def close(self):
        from rnvp_cabi import check, lib
        eng = self.module.engine()
        check(lib.rnvp_dp_finalize(eng.handle))
        eng.dp = None
