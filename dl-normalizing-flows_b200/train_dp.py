"""Data-parallel launcher with the semantics of the reference's ``train_flow`` (train.py:47-260)  --  SURVEY.md 8f-1/3.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 train_dp.py --datapath ... [--synthetic 6400]
    python train_dp.py ...                      (one GPU, no process group)

What it keeps from ``train_flow``: the data pipeline (ImageFolder -> Resize / CenterCrop / ToTensor, at most
100 batches, 90 / 10 train / validation split, train.py:65-100), ``RealNVP(channels, image_size, prior,
Hyperparameters(base_dim, res_blocks, True, True, True, True))`` (train.py:117-129), Adam(lr, weight_decay)
(train.py:134), resume from ``realnvp_state.pt`` / ``realnvp_state_optim.pt`` (train.py:139-154), the loss
``-mean(ll + logdet) + 5e-5 * weight_scale`` (train.py:192-194), the per-epoch bits/dim lines (train.py:203-207,
230-233), patience-100 early stopping (train.py:238-244), the checkpoint file names (train.py:249-250) and the 100
samples at the end (train.py:253-259).

What it changes (the "next" rows of SURVEY.md 8f):
  * one process per GPU: every global batch of ``--batch-size`` images is cut into equal per-rank shards; batch-norm
    statistics and gradients are synchronised by :mod:`rnvp_dp`, so the step equals a single process running the
    whole batch; rank 0 alone prints, checkpoints and samples;
  * the input pipeline is on the device: batches travel as uint8 (12 KB per image instead of 48 KB + log-det) and
    ``logit_transform`` (CPU in the reference, train.py:187) runs as a kernel; the running log-likelihood is
    accumulated on the device and read once per epoch instead of ``.item()`` every step (train.py:196);
  * the optimizer is the fused :class:`rnvp_optim.Adam`, whose ``state_dict`` has torch's layout: the checkpoint
    pair written here loads into the reference's ``train_flow --nofresh`` and vice versa;
  * validation runs the on-device bits/dim loop :func:`evaluate_bits_per_dim`, sharded over the ranks.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import torch
import torch.distributed as dist

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

SCALE_REG = 5e-5          # train.py:158
PATIENCE = 100            # train.py:243


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bits_per_dim(mean_logll: float, image_size: int, channels: int) -> float:
    """train.py:204, 231: (-mean_logll + ln(256) * D) / (D * ln 2)."""
    d = image_size * image_size * channels
    return (-mean_logll + math.log(256.0) * d) / (d * math.log(2.0))


def epoch_batches(n_items: int, batch_size: int, rank: int, world: int, seed: int, epoch: int, shuffle: bool = True):
    """Index tensors of this rank's shard of every global batch of one epoch.

    All ranks draw the same permutation (seeded by ``seed`` and ``epoch``), cut it into global batches of
    ``batch_size`` like ``DataLoader(shuffle=True)`` does, and take equal contiguous shards of each
    (``batch_size // world`` items; a ragged tail batch is trimmed to a multiple of ``world``, its remainder
    dropped) -- synchronised batch norm and the gradient average need equal local batches.
    """
    g = torch.Generator().manual_seed(seed * 1000003 + epoch)
    order = torch.randperm(n_items, generator=g) if shuffle else torch.arange(n_items)
    for b0 in range(0, n_items, batch_size):
        idx = order[b0:b0 + batch_size]
        per = idx.numel() // world
        if per == 0:
            continue
        yield idx[rank * per:(rank + 1) * per]


def load_images_uint8(datapath, dataset_name, image_size, channels, batch_size, synthetic, seed):
    """The (at most 100 batches of the) dataset as ONE uint8 tensor (N, C, H, W) plus the train / validation index
    split of train.py:81-86.  ``synthetic`` = N draws N smooth random images instead of reading an ImageFolder."""
    g = torch.Generator().manual_seed(seed)
    if synthetic:
        n = int(synthetic)
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, image_size), torch.linspace(-1, 1, image_size), indexing="ij")
        img = torch.zeros(n, channels, image_size, image_size)
        for _ in range(4):
            cx, cy = (torch.rand(n, channels, 1, 1, generator=g) * 2 - 1 for _ in range(2))
            sg = 0.15 + 0.5 * torch.rand(n, channels, 1, 1, generator=g)
            img += torch.rand(n, channels, 1, 1, generator=g) * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg ** 2))
        lo, hi = img.amin((1, 2, 3), keepdim=True), img.amax((1, 2, 3), keepdim=True)
        data = ((img - lo) / (hi - lo) * 255).round().to(torch.uint8)
    else:
        from torchvision import transforms
        from torchvision.datasets import ImageFolder
        tf = transforms.Compose([transforms.Resize((image_size, image_size)), transforms.CenterCrop(image_size),
                                 transforms.ToTensor()])
        ds = ImageFolder(root=os.path.join(datapath, dataset_name), transform=tf)
        keep = torch.randperm(len(ds), generator=g)[: batch_size * 100].tolist()          # train.py:81-82
        data = torch.stack([(ds[i][0] * 255).round().to(torch.uint8) for i in keep])       # ToTensor yields k / 255
    n = data.shape[0]
    perm = torch.randperm(n, generator=g)
    n_train = math.floor(n * 0.9)                                                          # train.py:84-86
    return data, perm[:n_train], perm[n_train:]


@torch.no_grad()
def evaluate_bits_per_dim(model, data_u8, indices, batch_size, image_size, channels, device, seed=0):
    """On-device validation loop (train.py:216-233): eval mode, mean over batches of the batch-mean log-likelihood
    (logit log-det included), sharded over the ranks; returns (mean_logll, bits/dim) on every rank."""
    import utils as U
    rank, world = rank_world()
    was_training = model.training
    model.eval()
    total = torch.zeros(2, dtype=torch.float64, device=device)          # (sum of per-batch means, batches)
    for shard in epoch_batches(indices.numel(), batch_size, rank, world, seed, 0, shuffle=False):
        x_u8 = data_u8[indices[shard]].to(device, non_blocking=True)
        x, logdet = U.logit_transform(x_u8)
        ll, _ = model(x)
        total[0] += (ll + logdet).mean().double()
        total[1] += 1
    if world > 1:
        # equal shards: the mean of the rank means of a batch is that batch's mean
        dist.all_reduce(total)
        total[1] /= world
        total[0] /= world
    model.train(was_training)
    mean_logll = float(total[0] / total[1].clamp(min=1))
    return mean_logll, bits_per_dim(mean_logll, image_size, channels)


def train_flow_dp(epochs, datapath, dataset_name, batch_size, image_size, channels, base_dim, res_blocks, output_dir,
                  fresh=True, saved_path=None, lr=5e-4, weight_decay=5e-5, synthetic=0, seed=999, math_mode="tf32",
                  max_batches=None, log=print):
    import flow_realnvp as F
    import rnvp_dp
    import rnvp_optim
    import utils as U
    rank, world = rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    say = log if rank == 0 else (lambda *a, **k: None)
    if batch_size % world:
        raise ValueError(f"--batch-size {batch_size} must be divisible by the number of ranks ({world})")

    torch.manual_seed(seed)                                             # main.py:58-59: identical initialisation everywhere
    data, train_idx, valid_idx = load_images_uint8(datapath, dataset_name, image_size, channels, batch_size, synthetic, seed)
    data = data.pin_memory()
    prior = torch.distributions.Normal(torch.tensor(0., device=device), torch.tensor(1., device=device), validate_args=False)
    net = F.RealNVP(channels, image_size, prior, U.Hyperparameters(base_dim, res_blocks, True, True, True, True)).to(device)
    net.set_math(math_mode)
    optimizer_state = None
    if not fresh:                                                       # train.py:139-154, same messages
        if saved_path is None:
            say("Fresh mode was disabled, but the '--saved-path' was not specified. See -h/--help for help.")
            return None
        try:
            net.load_state_dict(torch.load(os.path.join(saved_path, "realnvp_state.pt"), map_location=device))
            say("Loaded saved model.")
        except Exception:
            say("Could not load 'realnvp_state.pt' at {}, terminating.".format(os.path.join(saved_path, "realnvp_state.pt")))
            return None
        try:
            optimizer_state = torch.load(os.path.join(saved_path, "realnvp_state_optim.pt"), map_location=device)
        except Exception:
            say("Could not load 'realnvp_state_optim.pt' at {}, terminating.".format(os.path.join(saved_path, "realnvp_state_optim.pt")))
            return None
    model = rnvp_dp.DataParallel(net) if world > 1 else net
    optimizer = rnvp_optim.Adam(model, lr=lr, weight_decay=weight_decay)
    if optimizer_state is not None:
        optimizer.load_state_dict(optimizer_state)
        say("Loaded saved optimizer.")

    best, stale, epoch = float("-inf"), 0, 0
    history = []
    while epoch < epochs:
        epoch += 1
        say("Current epoch: {}".format(epoch))
        net.train()
        running = torch.zeros((), dtype=torch.float64, device=device)
        nb = 0
        for shard in epoch_batches(train_idx.numel(), batch_size, rank, world, seed, epoch):
            if max_batches is not None and nb >= max_batches:
                break
            x_u8 = data[train_idx[shard]].to(device, non_blocking=True)        # uint8 H2D
            optimizer.zero_grad()
            x, logdet = U.logit_transform(x_u8)                                # on the device
            ll, weight_scale = model(x)
            logll = (ll + logdet).mean()
            (-logll + SCALE_REG * weight_scale).backward()
            optimizer.step()
            running += logll.detach().double()
            nb += 1
        if world > 1:
            dist.all_reduce(running)
            running /= world
        mean_logll = float(running) / max(nb, 1)                               # the one host read of the epoch
        say("::Mean bits per dims: {}".format(bits_per_dim(mean_logll, image_size, channels)))
        v_logll, v_bpd = evaluate_bits_per_dim(model, data, valid_idx, batch_size, image_size, channels, device, seed)
        say("::Mean validation bits per dims: {}".format(v_bpd))
        history.append((bits_per_dim(mean_logll, image_size, channels), v_bpd))
        if v_logll > best:
            best, stale = v_logll, 0
        else:
            stale += 1
            if stale >= PATIENCE:
                break
    say("Training finished at epoch {} with log-likelihood {}".format(epoch, best))
    if rank == 0:
        os.makedirs(os.path.join(output_dir, "states"), exist_ok=True)
        os.makedirs(os.path.join(output_dir, "gen"), exist_ok=True)
        torch.save(net.state_dict(), os.path.join(output_dir, "states", "realnvp_state.pt"))                # train.py:249
        torch.save(optimizer.state_dict(), os.path.join(output_dir, "states", "realnvp_state_optim.pt"))    # train.py:250
        net.eval()
        with torch.no_grad():
            imgs, _ = U.logit_transform(net.sample(size=100), reverse=True)
        try:
            import torchvision
            torchvision.utils.save_image(imgs, os.path.join(output_dir, "gen", "img_realnvp.png"), nrow=10)
        except Exception as e:                                       # pragma: no cover - torchvision / PIL optional
            torch.save(imgs.cpu(), os.path.join(output_dir, "gen", "img_realnvp.pt"))
            say(f"(torchvision.utils.save_image unavailable: {e}; samples saved as a tensor)")
    if world > 1:
        dist.barrier()
        model.close()
    return history


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    # the RealNVP flags of the reference's CLI (main.py:123-267), same names and defaults
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--learning-rate", type=float, default=5e-4)
    ap.add_argument("--weight-decay", type=float, default=5e-5)
    ap.add_argument("--dataset-name", default="kaggle-full")
    ap.add_argument("--datapath", default="./datasets")
    ap.add_argument("--batch-size", type=int, default=64, help="GLOBAL batch (split evenly over the ranks)")
    ap.add_argument("--image-size", type=int, default=64)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--output-dir", default="./output")
    ap.add_argument("--nofresh", action="store_true")
    ap.add_argument("--saved-path", default=None)
    ap.add_argument("--fixed-seed", type=int, default=999)
    ap.add_argument("--base-dim", type=int, default=64)
    ap.add_argument("--res-blocks", type=int, default=8)
    # additions
    ap.add_argument("--synthetic", type=int, default=0, help="train on N smooth synthetic images instead of an ImageFolder")
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--max-batches", type=int, default=None, help="cap the batches per epoch (smoke runs)")
    a = ap.parse_args(argv)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    try:
        return train_flow_dp(a.epochs, a.datapath, a.dataset_name, a.batch_size, a.image_size, a.channels, a.base_dim,
                             a.res_blocks, a.output_dir, fresh=not a.nofresh, saved_path=a.saved_path,
                             lr=a.learning_rate, weight_decay=a.weight_decay, synthetic=a.synthetic, seed=a.fixed_seed,
                             math_mode=a.math, max_batches=a.max_batches)
    finally:
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
