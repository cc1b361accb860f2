"""CPU, world_size 2, gloo: the frame-sharding protocol (two all-reduces per pass) gives the
same correlation / mean-distance sums as one rank.  The kernels are mocked with numpy here
(no GPU in this container); what is under test is the host logic of sharding.py that
FramePipeline runs on NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wisp_mdanalysis_implementation_b200 import sharding


def _local_pass(nodes, lo, hi, group=None):
    """numpy stand-in for K1..K3 on frames [lo, hi) + the real sharding collectives."""
    T, N, _ = nodes.shape
    x = nodes[lo:hi].reshape(hi - lo, 3 * N)
    colsum = torch.from_numpy(x.sum(axis=0))
    mean = torch.empty_like(colsum)
    sharding.global_mean(colsum, mean, T, group=group)
    xc = x - mean.numpy()
    g = np.einsum("tic,tjc->ij", xc.reshape(-1, N, 3), xc.reshape(-1, N, 3))
    pos = x.reshape(-1, N, 3)
    d = np.sqrt(((pos[:, :, None, :] - pos[:, None, :, :]) ** 2).sum(-1)).sum(axis=0)
    sums = torch.from_numpy(np.stack([g, d]))
    sharding.reduce_sums(sums, group=group)
    return sums.numpy(), mean.numpy()


def _worker(rank, world, port, nodes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.frame_range(nodes.shape[0], rank, world)
    sums, mean = _local_pass(nodes, lo, hi)
    np.save(os.path.join(out_dir, "sums_%d.npy" % rank), sums)
    np.save(os.path.join(out_dir, "mean_%d.npy" % rank), mean)
    dist.destroy_process_group()


def test_frame_range_partitions():
    for T in (1, 7, 50000, 50001):
        for world in (1, 2, 3, 8):
            blocks = [sharding.frame_range(T, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == T
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_reduction_matches_single_rank(tmp_path):
    rng = np.random.default_rng(0)
    nodes = rng.normal(size=(41, 9, 3)) + rng.normal(scale=10, size=(1, 9, 3))
    want, want_mean = _local_pass(nodes, 0, nodes.shape[0])
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, nodes, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / ("sums_%d.npy" % r))
        mean = np.load(tmp_path / ("mean_%d.npy" % r))
        np.testing.assert_allclose(mean, want_mean, rtol=1e-13)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)
