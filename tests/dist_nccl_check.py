"""Run under torchrun (one rank per GPU, NCCL): the sharded gallery and the row-sharded fit must reproduce the
single-GPU results bit for bit.  Launched by tests/test_gpu_dist.py when >= 2 GPUs are visible.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_nccl_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    run_checks(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print(f"dist_nccl_check ok (world {world})", flush=True)
    dist.destroy_process_group()


def run_checks(rank, world, dev):
    """The assertions proper; needs an initialised NCCL process group (bench.py --gpus N calls this too, so that the
    driver's scaling run proves them on its own box)."""
    rng = np.random.default_rng(3)                        # identical data on every rank
    # ---- sharded gallery (BASELINE config 3 in miniature)
    n, k, B = 200_003, 128, 512
    lam = 1.0 / np.arange(1, k + 1) ** 2
    G = rng.normal(size=(n, k)) * np.sqrt(lam)
    G[150_000] = G[17]                                    # duplicate in another shard
    truth = rng.integers(0, n, B); truth[0] = 17
    P = G[truth] + 0.05 * rng.normal(size=(B, k)) * np.sqrt(lam); P[0] = G[17]
    p = torch.from_numpy(P).to(dev)
    for metric in (ef.METRIC_COSINE_SK, ef.METRIC_COSINE_G1, ef.METRIC_L2):
        lo, hi = ef.dist.shard_bounds(n, world, rank)
        shard = ef.dist.ShardedGallery(G[lo:hi], lo, metric)
        s, i = shard.match(p)
        whole = ef.dist.ShardedGallery(G, 0, metric, group=dist.new_group([rank]) if False else None)
        ws, wi = whole.match_local(p)                     # single-GPU answer computed on every rank
        assert torch.equal(i, wi), f"metric {metric}: sharded argbest differs from the unsharded one"
        assert torch.allclose(s, ws, rtol=1e-13, atol=0)
        assert int(i[0]) == 17
        if metric != ef.METRIC_L2 and rank == 0:
            acc = float((i.cpu().numpy() == truth).mean())
            assert acc > 0.95, acc
    # ---- row-sharded fit (covariance branch)
    N, D, kk = 6000, 256, 16
    base = rng.normal(size=(N, 12)) @ rng.normal(size=(12, D))
    X = np.clip(np.rint(128 + 20 * base + rng.normal(0, 4, (N, D))), 0, 255).astype(np.uint8)
    lo, hi = ef.dist.shard_bounds(N, world, rank)
    E, mean, proj, ev = ef.dist.fit_gen1_sharded(torch.from_numpy(X[lo:hi]).to(dev), N, kk)
    E1, mean1, proj1, ev1, _ = ef.fit_gen1(X, kk)         # single-GPU engine fit of the whole matrix
    np.testing.assert_allclose(ev.cpu().numpy(), ev1, rtol=1e-9)
    assert np.array_equal(mean.cpu().numpy(), mean1)
    sign = np.sign(np.sum(E.cpu().numpy() * E1, axis=0))
    np.testing.assert_allclose(E.cpu().numpy() * sign, E1, atol=1e-8)
    np.testing.assert_allclose(proj.cpu().numpy() * sign, proj1[lo:hi], atol=1e-6)
    # ---- subspace solver with the covariance products sharded over the ranks: bit identical to the unsharded solve
    solo = [dist.new_group([r]) for r in range(world)][rank]      # every rank creates every group, keeps its own
    Mx = rng.normal(size=(900, 900)); Cm = torch.from_numpy(Mx @ Mx.T / 900 + np.diag(np.linspace(50, 0, 900))).to(dev)
    lam_s, Q_s, info_s = ef.dist.eigh_topk_device(Cm, 20)                       # sharded over the default group
    lam_1, Q_1, info_1 = ef.dist.eigh_topk_device(Cm, 20, group=solo)           # same solve on one rank
    assert info_s.get("allgathers", 0) > 0 and "allgathers" not in info_1
    assert torch.equal(lam_s, lam_1) and torch.equal(Q_s, Q_1), "sharded covariance products changed the iterates"
    w_ref = np.linalg.eigvalsh(Cm.cpu().numpy())[::-1][:20]
    np.testing.assert_allclose(lam_s.cpu().numpy(), w_ref, rtol=1e-9)
    Es, means, projs, evs = ef.dist.fit_gen1_sharded(torch.from_numpy(X[lo:hi]).to(dev), N, kk, solver="subspace")
    np.testing.assert_allclose(evs.cpu().numpy(), ev1, rtol=1e-9)
    gathered = [torch.empty_like(ev) for _ in range(world)]
    dist.all_gather(gathered, ev)
    assert all(torch.equal(g, ev) for g in gathered), "eigenvalues must be bit identical on every rank"


if __name__ == "__main__":
    main()
