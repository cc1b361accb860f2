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

from wisp_mdanalysis_implementation_b200 import paths_sharded, sharding


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


def _fake_query(s, t):
    """Stand-in for wisp_paths_query: a result that depends on the pair only."""
    return [[0.25 * (s + t) + k, s] + list(range(s + 1, s + 1 + k)) + [t] for k in range(1 + (s + t) % 3)]


def _pairs_worker(rank, world, port, pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = paths_sharded.query_pairs(None, None, pairs, 3, query=_fake_query)
    if rank == 0:
        import pickle
        with open(os.path.join(out_dir, "pairs.pkl"), "wb") as f:
            pickle.dump(got, f)
    else:
        assert got is None
    dist.destroy_process_group()


def test_pair_sharding_covers_every_pair_once():
    for n, world in ((0, 2), (1, 4), (7, 2), (200, 8)):
        seen = sorted(i for r in range(world) for i in paths_sharded.pair_shard(n, r, world))
        assert seen == list(range(n))


def test_two_rank_pair_queries_gather_in_order(tmp_path):
    import pickle
    pairs = [(3, 11), (0, 5), (7, 2), (4, 4 + 9), (1, 30), (6, 8), (2, 19)]
    want = [_fake_query(s, t) for s, t in pairs]
    assert paths_sharded.query_pairs(None, None, pairs, 3, query=_fake_query) == want   # no process group
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_pairs_worker, args=(2, port, pairs, str(tmp_path)), nprocs=2, join=True)
    with open(tmp_path / "pairs.pkl", "rb") as f:
        assert pickle.load(f) == want


def _streamed_pass(nodes, lo, hi, rank, n_blocks=2):
    """numpy stand-in for FramePipeline.run_host_streamed: blocks centred with a provisional
    reference (mean of rank 0's first block, broadcast), ONE all-reduce of [G || D || s], exact
    parallel-axis correction afterwards (what K4 does with d_offset_sum)."""
    T, N, _ = nodes.shape
    x = nodes[lo:hi].reshape(hi - lo, 3 * N)
    bounds = np.linspace(0, hi - lo, n_blocks + 1).astype(int)
    ref = torch.from_numpy(x[bounds[0]:bounds[1]].mean(axis=0).copy())
    if sharding.is_distributed():
        dist.broadcast(ref, src=0)
    ref = ref.numpy()
    g = np.zeros((N, N))
    d = np.zeros((N, N))
    for b in range(n_blocks):
        xb = x[bounds[b]:bounds[b + 1]] - ref
        g += np.einsum("tic,tjc->ij", xb.reshape(-1, N, 3), xb.reshape(-1, N, 3))
        pos = x[bounds[b]:bounds[b + 1]].reshape(-1, N, 3)        # distances do not care about the reference
        d += np.sqrt(((pos[:, :, None, :] - pos[:, None, :, :]) ** 2).sum(-1)).sum(axis=0)
    s = (x - ref).sum(axis=0)
    flat = torch.from_numpy(np.concatenate([g.ravel(), d.ravel(), s]))
    sharding.reduce_sums(flat)
    flat = flat.numpy()
    g, d, s = flat[:N * N].reshape(N, N), flat[N * N:2 * N * N].reshape(N, N), flat[2 * N * N:].reshape(N, 3)
    return g - (s @ s.T) / T, d, ref + s.reshape(-1) / T


def _streamed_worker(rank, world, port, nodes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.frame_range(nodes.shape[0], rank, world)
    g, d, mean = _streamed_pass(nodes, lo, hi, rank)
    np.savez(os.path.join(out_dir, "streamed_%d.npz" % rank), g=g, d=d, mean=mean)
    dist.destroy_process_group()


def test_two_rank_streamed_protocol_matches_centred_gram(tmp_path):
    """The host-fed mode's single-collective protocol: provisional reference + parallel-axis
    correction gives the mean-centred Gram, the distance sums and the mean of one rank."""
    rng = np.random.default_rng(3)
    nodes = rng.normal(size=(53, 7, 3)) + rng.normal(scale=20, size=(1, 7, 3))
    T, N, _ = nodes.shape
    xc = nodes - nodes.mean(axis=0)
    want_g = np.einsum("tic,tjc->ij", xc, xc)
    want_d = np.sqrt(((nodes[:, :, None, :] - nodes[:, None, :, :]) ** 2).sum(-1)).sum(axis=0)
    g1, d1, m1 = _streamed_pass(nodes, 0, T, 0)                       # no process group
    np.testing.assert_allclose(g1, want_g, rtol=1e-10, atol=1e-9)
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_streamed_worker, args=(2, port, nodes, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / ("streamed_%d.npz" % r))
        np.testing.assert_allclose(got["g"], want_g, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(got["d"], want_d, rtol=1e-12)
        np.testing.assert_allclose(got["mean"], nodes.mean(axis=0).reshape(-1), rtol=1e-13)
