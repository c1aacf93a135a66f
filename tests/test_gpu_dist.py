"""Multi-GPU paths on real devices: world size 1 in-process, and NCCL with one rank per GPU when >= 2 GPUs exist."""
import os
import subprocess
import sys

import numpy as np
import pytest

import eigenfaces_b200 as ef
from conftest import ROOT
from gpu_util import require_gpu
from oracle import gen1, gen2

pytestmark = pytest.mark.gpu


def test_sharded_gallery_single_rank_matches_oracle():
    torch = require_gpu()
    rng = np.random.default_rng(1)
    n, k, B = 30_000, 128, 300
    G = rng.normal(size=(n, k)) / np.arange(1, k + 1)
    P = G[rng.integers(0, n, B)] + 0.02 * rng.normal(size=(B, k)) / np.arange(1, k + 1)
    p = torch.from_numpy(P).cuda()
    # three shards processed in one process, reduced with the same kernels the NCCL path uses
    for metric in (ef.METRIC_COSINE_SK, ef.METRIC_L2):
        scores, idxs = [], []
        for r in range(3):
            lo, hi = ef.dist.shard_bounds(n, 3, r)
            s, i = ef.dist.ShardedGallery(G[lo:hi], lo, metric).match_local(p)
            scores.append(s); idxs.append(i)
        bs, bi = ef.dist.reduce_candidates(torch.stack(scores), torch.stack(idxs), metric)
        if metric == ef.METRIC_L2:
            want = np.array([((G - q) ** 2).sum(1).argmin() for q in P])
        else:
            want = gen2.sk_cosine_similarity(P, G).argmax(1)
        assert np.array_equal(bi.cpu().numpy(), want)


def test_row_sharded_fit_world1_matches_oracle():
    torch = require_gpu()
    rng = np.random.default_rng(2)
    N, D, k = 3000, 192, 12
    base = rng.normal(size=(N, 10)) @ rng.normal(size=(10, D))
    X = np.clip(np.rint(128 + 20 * base + rng.normal(0, 4, (N, D))), 0, 255).astype(np.uint8)
    E, mean, proj, ev = ef.dist.fit_gen1_sharded(torch.from_numpy(X).cuda(), N, k)
    E_ref, mean_ref, proj_ref, ev_ref = gen1.manual_pca(X.astype(np.float64), k)
    np.testing.assert_allclose(ev.cpu().numpy(), ev_ref, rtol=1e-9)
    np.testing.assert_allclose(mean.cpu().numpy(), mean_ref, rtol=1e-15)
    sign = np.sign(np.sum(E.cpu().numpy() * E_ref, axis=0))
    np.testing.assert_allclose(E.cpu().numpy() * sign, E_ref, atol=1e-8)
    np.testing.assert_allclose(proj.cpu().numpy() * sign, proj_ref, atol=1e-6)


def test_two_devices_one_process():
    """One process driving two GPUs: the per-device cudaFuncSetAttribute bookkeeping (EF_ENSURE_SMEM) must give the
    second device its shared-memory attributes too; the same model on cuda:0 and cuda:1 returns identical results
    through the serving kernel, the cluster kernel and K1."""
    torch = require_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    rng = np.random.default_rng(3)
    D, k, ng, B = 4096, 10, 300, 512
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    mean = rng.uniform(60, 200, D)
    gal = rng.normal(size=(ng, k)) * 100
    x = rng.integers(0, 256, (B, D), dtype=np.uint8)
    frames = rng.integers(0, 256, (2, 240, 320), dtype=np.uint8)
    boxes = np.array([[i % 2, 5 + i, 7 + i, 100 + i, 90 + i] for i in range(40)], dtype=np.int32)
    got = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            rec = ef.Recognizer(E, mean, gal, metric=ef.METRIC_COSINE_G1, labels=np.arange(ng) % 4)
            xd = torch.from_numpy(x).cuda()
            one = rec.recognize_device(xd, 0.5)
            outs = [rec.submit_device(xd, 0.5) for _ in range(3)]
            rec.flush_device()
            crops = ef.engine.preprocess_device(torch.from_numpy(frames).cuda(), torch.from_numpy(boxes).cuda(), 64)
            torch.cuda.synchronize()
            assert rec.pipeline_timeouts() == 0
            for o in outs:
                for f in ("score", "index", "label", "resid2"):
                    assert torch.equal(o[f], one[f]), f
            got.append(({f: one[f].cpu() for f in ("score", "index", "label", "resid2")}, crops.cpu()))
            rec.close()
    for f in got[0][0]:
        assert torch.equal(got[0][0][f], got[1][0][f]), f
    assert torch.equal(got[0][1], got[1][1])


def test_nccl_two_ranks_bit_identical():
    torch = require_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_nccl_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "dist_nccl_check ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
