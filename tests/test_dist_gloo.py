"""Host-side logic of the multi-GPU paths under a real 2-rank gloo process group on CPU: shard bounds, the
(score, index) all-gather + tie-aware reduction of the sharded gallery, and the exact integer all-reduce of the
row-sharded Gram.  The per-shard arithmetic that runs in CUDA kernels on a GPU box is supplied here by the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import eigenfaces_b200 as ef
from oracle import gen1, gen2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            b = [ef.dist.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _sharded_match(rank, world):
    rng = np.random.default_rng(7)                         # same data on every rank
    n, k, B = 1001, 12, 40
    G = rng.normal(size=(n, k))
    G[700] = G[3]; G[1000] = G[3]                          # exact duplicates across shards
    P = np.concatenate([G[rng.integers(0, n, B - 1)] + 0.01 * rng.normal(size=(B - 1, k)), G[3:4]])
    lo, hi = ef.dist.shard_bounds(n, world, rank)
    res = {}
    for metric in (ef.METRIC_COSINE_SK, ef.METRIC_L2):
        if metric == ef.METRIC_L2:
            d = ((P[:, None, :] - G[None, lo:hi]) ** 2).sum(-1)
            score, idx = d.min(1), d.argmin(1) + lo
        else:
            s = gen2.sk_cosine_similarity(P, G[lo:hi])
            score, idx = s.max(1), s.argmax(1) + lo
        scores, idxs = ef.dist.allgather_candidates(torch.from_numpy(score), torch.from_numpy(idx))
        bs, bi = ef.dist.reduce_candidates(scores, idxs, metric)
        res[metric] = (bs.numpy(), bi.numpy())
    return res, P, G


def test_sharded_gallery_equals_unsharded_argmax_gloo():
    results = _run(_sharded_match, 2)
    (r0, P, G), (r1, _, _) = results
    for metric in (ef.METRIC_COSINE_SK, ef.METRIC_L2):
        assert np.array_equal(r0[metric][1], r1[metric][1]) and np.array_equal(r0[metric][0], r1[metric][0])
        if metric == ef.METRIC_L2:
            want = ((P[:, None, :] - G[None]) ** 2).sum(-1).argmin(1)
        else:
            want = gen2.sk_cosine_similarity(P, G).argmax(1)
        assert np.array_equal(r0[metric][1], want)
        assert r0[metric][1][-1] == 3                      # three identical rows in two shards: lowest global row


def test_reduce_candidates_rules():
    s = torch.tensor([[0.5, 0.9, -1.0], [0.5, 0.8, -1.0], [0.7, 0.9, 0.0]], dtype=torch.float64)
    i = torch.tensor([[10, 11, 12], [4, 5, -1], [20, 2, -1]], dtype=torch.int64)
    bs, bi = ef.dist.reduce_candidates(s, i, ef.METRIC_COSINE_SK)
    assert bi.tolist() == [20, 2, 12] and bs.tolist() == [0.7, 0.9, -1.0]
    bs, bi = ef.dist.reduce_candidates(s, i, ef.METRIC_L2)
    assert bi.tolist() == [4, 5, 12]
    empty = torch.full((2, 2), -1, dtype=torch.int64)
    assert ef.dist.reduce_candidates(torch.zeros(2, 2, dtype=torch.float64), empty, ef.METRIC_L2)[1].tolist() == [-1, -1]


def _sharded_gram(rank, world):
    rng = np.random.default_rng(11)
    N, D = 301, 48
    X = rng.integers(0, 256, (N, D), dtype=np.uint8)
    lo, hi = ef.dist.shard_bounds(N, world, rank)
    Xi = X[lo:hi].astype(np.int64)
    buf = torch.from_numpy(np.concatenate([(Xi.T @ Xi).ravel(), Xi.sum(0)]))   # what ef_gram_u8 / ef_colsum_u8 produce
    ef.dist.allreduce_exact(buf)
    G, s = buf[:D * D].numpy().reshape(D, D), buf[D * D:].numpy()
    cov = (N * G - np.outer(s, s)).astype(np.float64) / N / (N - 1)             # ef_gram_center_device, side 1
    return cov, X


def test_row_sharded_gram_allreduce_gloo():
    (c0, X), (c1, _) = _run(_sharded_gram, 2)
    assert np.array_equal(c0, c1)                                              # exact, order independent
    ref = np.cov((X.astype(np.float64) - X.mean(0)).T)
    np.testing.assert_allclose(c0, ref, rtol=1e-11, atol=1e-9)
    w = np.linalg.eigvalsh(c0)[::-1][:5]
    np.testing.assert_allclose(w, gen1.manual_pca(X.astype(np.float64), 5)[3], rtol=1e-10)
