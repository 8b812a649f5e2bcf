"""The data-parallel launcher (train_dp.py): host logic on the CPU (gloo), the loop itself on a GPU."""
import math
import os
import sys

import pytest
import torch

from conftest import PKG_DIR

sys.path.insert(0, PKG_DIR)


def test_epoch_batches_partition_the_global_batches():
    import train_dp as T
    n, bs, world = 103, 16, 4
    per_rank = [list(T.epoch_batches(n, bs, r, world, seed=7, epoch=3)) for r in range(world)]
    nb = len(per_rank[0])
    assert all(len(p) == nb for p in per_rank) and nb == math.ceil(n / bs)
    seen = []
    for b in range(nb):
        shards = [per_rank[r][b] for r in range(world)]
        assert len({s.numel() for s in shards}) == 1                  # equal local batches (sync BN, ncclAvg)
        seen += torch.cat(shards).tolist()
    assert len(seen) == len(set(seen))                                # nothing is used twice
    assert len(seen) == (n // bs) * bs + ((n % bs) // world) * world  # ragged tail trimmed to a multiple of world
    # a different epoch reshuffles, the same epoch is reproducible on every rank
    again = list(T.epoch_batches(n, bs, 1, world, seed=7, epoch=3))
    assert all(torch.equal(a, b) for a, b in zip(again, per_rank[1]))
    other = list(T.epoch_batches(n, bs, 1, world, seed=7, epoch=4))
    assert not all(torch.equal(a, b) for a, b in zip(other, per_rank[1]))


def test_bits_per_dim_formula():
    import train_dp as T
    # etc/realnvp_logll.csv:2 <-> etc/realnvp_bpd.csv:26 of the reference: 35757.12 nats <-> 3.8019 bits/dim at 64x64x3
    assert abs(T.bits_per_dim(35757.12, 64, 3) - 3.8019) < 5e-4


@pytest.mark.gpu
def test_train_flow_dp_single_gpu(tmp_path, capsys):
    """Two short epochs on synthetic images: the log lines of train.py, a falling bits/dim, the checkpoint pair under
    the reference's file names in the reference's layout, and a resume (--nofresh) that continues from it."""
    import train_dp as T
    out = str(tmp_path)
    common = dict(datapath=".", dataset_name="none", batch_size=16, image_size=32, channels=3, base_dim=8, res_blocks=1,
                  output_dir=out, synthetic=160, seed=999, math_mode="tf32")
    hist = T.train_flow_dp(epochs=2, **common)
    text = capsys.readouterr().out
    for line in ("Current epoch: 1", "::Mean bits per dims:", "::Mean validation bits per dims:", "Training finished at epoch 2"):
        assert line in text, text
    assert len(hist) == 2 and all(math.isfinite(v) for pair in hist for v in pair)
    assert hist[1][0] < hist[0][0]                                    # it trains
    state = torch.load(os.path.join(out, "states", "realnvp_state.pt"), map_location="cpu")
    opt_state = torch.load(os.path.join(out, "states", "realnvp_state_optim.pt"), map_location="cpu")
    import realnvp_oracle as O
    assert [(k, tuple(v.shape)) for k, v in state.items()] == list(O.state_shapes(3, 32, 8, 1, 5).items())
    assert set(opt_state) == {"state", "param_groups"} and opt_state["param_groups"][0]["lr"] == 5e-4
    assert os.path.exists(os.path.join(out, "gen", "img_realnvp.png")) or os.path.exists(os.path.join(out, "gen", "img_realnvp.pt"))
    hist2 = T.train_flow_dp(epochs=1, fresh=False, saved_path=os.path.join(out, "states"), **common)
    text = capsys.readouterr().out
    assert "Loaded saved model." in text and "Loaded saved optimizer." in text
    assert hist2[0][0] < hist[0][0]                                   # continues from the trained state
    assert T.train_flow_dp(epochs=1, fresh=False, saved_path=None, **common) is None      # train.py:140-142
