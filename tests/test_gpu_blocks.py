"""Device building blocks of the fit and of the sharded gallery, each against numpy."""
import ctypes as C

import numpy as np
import pytest

import eigenfaces_b200 as ef
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_dgemm_strided_layouts():
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(0)
    for (M, N, K) in ((70, 33, 129), (256, 256, 64), (5, 300, 1000), (1, 1, 1)):
        A = rng.normal(size=(M, K)); B = rng.normal(size=(K, N)); C0 = rng.normal(size=(M, N))
        for ta in (False, True):
            for tb in (False, True):
                a = torch.from_numpy(np.ascontiguousarray(A.T if ta else A)).cuda()
                b = torch.from_numpy(np.ascontiguousarray(B.T if tb else B)).cuda()
                c = torch.from_numpy(C0.copy()).cuda()
                sam, sak = (1, M) if ta else (K, 1)
                sbk, sbn = (1, K) if tb else (N, 1)
                ef._lib.check(L.ef_dgemm_device(M, N, K, 0.5, a.data_ptr(), sam, sak, b.data_ptr(), sbk, sbn, 2.0,
                                                c.data_ptr(), N, _stream(torch)), "dgemm")
                np.testing.assert_allclose(c.cpu().numpy(), 0.5 * A @ B + 2.0 * C0, rtol=1e-12, atol=1e-12)


def test_dgemm_tensor_core_layouts_edges_and_split_k():
    """ef_dgemm_tc_device (mma.sync m8n8k4 f64) against numpy: all four operand layouts, ragged M / N / K, odd leading
    strides (8-byte copies), split-K with its ordered reduction; sharding the rows of A must not change a single bit."""
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(1)
    for (M, N, K, splits) in ((70, 33, 129, 1), (256, 256, 64, 1), (5, 300, 1000, 4), (1, 1, 1, 1), (320, 320, 4099, 16),
                              (1000, 130, 17, 1), (129, 257, 3, 1)):
        A = rng.normal(size=(M, K)); B = rng.normal(size=(K, N)); C0 = rng.normal(size=(M, N))
        for ta in (False, True):
            for tb in (False, True):
                a = torch.from_numpy(np.ascontiguousarray(A.T if ta else A)).cuda()
                b = torch.from_numpy(np.ascontiguousarray(B.T if tb else B)).cuda()
                c = torch.from_numpy(C0.copy()).cuda()
                sam, sak = (1, M) if ta else (K, 1)
                sbk, sbn = (1, K) if tb else (N, 1)
                wb = int(L.ef_dgemm_tc_work_bytes(M, N, splits))
                work = torch.empty(max(wb, 16), dtype=torch.uint8, device="cuda")
                ef._lib.check(L.ef_dgemm_tc_device(M, N, K, 0.5, a.data_ptr(), sam, sak, b.data_ptr(), sbk, sbn, 2.0,
                                                   c.data_ptr(), N, splits, work.data_ptr(), _stream(torch)), "dgemm_tc")
                np.testing.assert_allclose(c.cpu().numpy(), 0.5 * A @ B + 2.0 * C0, rtol=1e-12, atol=1e-11)
    # rows of A sharded 3 ways: every output element is computed in the same order whoever owns its row
    M, N, K = 1000, 200, 777
    a = torch.from_numpy(rng.normal(size=(M, K))).cuda(); b = torch.from_numpy(rng.normal(size=(K, N))).cuda()
    whole = torch.empty((M, N), dtype=torch.float64, device="cuda")
    parts = torch.empty((M, N), dtype=torch.float64, device="cuda")
    ef._lib.check(L.ef_dgemm_tc_device(M, N, K, 1.0, a.data_ptr(), K, 1, b.data_ptr(), N, 1, 0.0, whole.data_ptr(), N, 1, None,
                                       _stream(torch)), "dgemm_tc")
    for lo, hi in ((0, 333), (333, 334), (334, 1000)):
        ef._lib.check(L.ef_dgemm_tc_device(hi - lo, N, K, 1.0, a[lo:hi].data_ptr(), K, 1, b.data_ptr(), N, 1, 0.0,
                                           parts[lo:hi].data_ptr(), N, 1, None, _stream(torch)), "dgemm_tc")
    assert torch.equal(whole, parts)


def test_cholesky_inverse_kernel():
    """ef_chol_inverse_device: G = L L^T and L^-1 against numpy, ragged sizes; a non-positive pivot is reported."""
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(2)
    for m in (1, 7, 32, 33, 100, 320, 417):
        Y = rng.normal(size=(m + 50, m))
        G = Y.T @ Y
        g = torch.from_numpy(G.copy()).cuda()
        linv = torch.empty((m, m), dtype=torch.float64, device="cuda")
        info = torch.full((1,), -1, dtype=torch.int32, device="cuda")
        ef._lib.check(L.ef_chol_inverse_device(g.data_ptr(), m, linv.data_ptr(), info.data_ptr(), _stream(torch)), "chol")
        assert int(info) == 0
        Lref = np.linalg.cholesky(G)
        np.testing.assert_allclose(np.tril(g.cpu().numpy()), Lref, rtol=1e-10, atol=1e-10)
        Li = linv.cpu().numpy()
        assert np.abs(np.triu(Li, 1)).max() == 0.0
        np.testing.assert_allclose(Li @ G @ Li.T, np.eye(m), atol=1e-9)
    bad = np.eye(40); bad[17, 17] = -1.0
    g = torch.from_numpy(bad).cuda()
    linv = torch.empty((40, 40), dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    ef._lib.check(L.ef_chol_inverse_device(g.data_ptr(), 40, linv.data_ptr(), info.data_ptr(), _stream(torch)), "chol")
    assert int(info) == 18


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 229, 300, 321, 512, 590, 640, 700])
def test_jacobi_eigensolver(n):
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(n)
    Z = rng.normal(size=(n, max(n // 2, 1) if n > 8 else n + 3))        # rank deficient for n > 8 (like a centred Gram)
    A = Z @ Z.T
    a = torch.from_numpy(A.copy()).cuda()
    evals = torch.empty(n, dtype=torch.float64, device="cuda")
    evecs = torch.empty((n, n), dtype=torch.float64, device="cuda")
    work = torch.empty(L.ef_eigh_work_bytes(n), dtype=torch.uint8, device="cuda")
    sweeps, off = C.c_int32(), C.c_double()
    ef._lib.check(L.ef_eigh_jacobi_device(a.data_ptr(), n, evals.data_ptr(), evecs.data_ptr(), work.data_ptr(), 0, 0.0,
                                          C.byref(sweeps), C.byref(off), _stream(torch)), "jacobi")
    w = evals.cpu().numpy(); V = evecs.cpu().numpy()
    w_ref = np.linalg.eigvalsh(A)[::-1]
    np.testing.assert_allclose(w, w_ref, rtol=1e-10, atol=1e-10 * max(1.0, abs(w_ref[0])))
    assert np.all(np.diff(w) <= 1e-9 * max(1.0, abs(w[0])))                    # descending
    np.testing.assert_allclose(V @ V.T, np.eye(n), atol=1e-11)
    np.testing.assert_allclose(V @ A @ V.T, np.diag(w), atol=1e-9 * max(1.0, abs(w_ref[0])))
    assert sweeps.value <= (20 if n <= 320 else 32)        # the blocked ordering (n > 320) needs more sweeps on a rank-deficient matrix


def test_integer_gram_colsum_and_centring():
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(1)
    N, D = 150, 1000
    X = rng.integers(0, 256, (N, D), dtype=np.uint8)
    x = torch.from_numpy(X).cuda()
    cs = torch.empty(D, dtype=torch.int64, device="cuda")
    ef._lib.check(L.ef_colsum_u8_device(x.data_ptr(), D, N, D, cs.data_ptr(), _stream(torch)), "colsum")
    assert np.array_equal(cs.cpu().numpy(), X.astype(np.int64).sum(0))
    Xi = X.astype(np.int64)
    for side, n in ((0, N), (1, D)):
        G = torch.zeros((n, n), dtype=torch.int64, device="cuda")
        # accumulate in two column (side 0) / row (side 1) chunks: the += contract of the sharded fit
        if side == 0:
            for d0, d1 in ((0, 333), (333, D)):
                ef._lib.check(L.ef_gram_u8_device(x.data_ptr(), D, N, D, d0, d1, 0, G.data_ptr(), _stream(torch)), "gram")
            want = Xi @ Xi.T
        else:
            for r0, r1 in ((0, 70), (70, N)):
                ef._lib.check(L.ef_gram_u8_device(x[r0:r1].data_ptr(), D, r1 - r0, D, 0, D, 1, G.data_ptr(), _stream(torch)), "gram")
            want = Xi.T @ Xi
        assert np.array_equal(G.cpu().numpy(), want)
        Cc = torch.empty((n, n), dtype=torch.float64, device="cuda")
        work = torch.empty(max(L.ef_gram_center_work_bytes(n), 16), dtype=torch.uint8, device="cuda")
        ef._lib.check(L.ef_gram_center_device(G.data_ptr(), n, side, cs.data_ptr(), N, 1.0 / (N - 1), Cc.data_ptr(),
                                              work.data_ptr(), _stream(torch)), "center")
        Xc = X.astype(np.float64) - X.astype(np.float64).mean(0)
        ref = (Xc @ Xc.T if side == 0 else Xc.T @ Xc) / (N - 1)
        np.testing.assert_allclose(Cc.cpu().numpy(), ref, rtol=1e-10, atol=1e-8)


def test_match_kernels_sharded_equals_whole():
    """Sharded gallery: per-shard top-1 + reduce == whole-gallery argmax (ties -> lowest global index)."""
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(7)
    n, k, B = 5000, 24, 77
    G = rng.normal(size=(n, k)); G[4000] = G[100]; G[4999] = G[100]       # duplicates across shards
    P = np.concatenate([G[rng.integers(0, n, B - 1)] + 0.01 * rng.normal(size=(B - 1, k)), G[100:101]])
    g = torch.from_numpy(G).cuda(); p = torch.from_numpy(P).cuda()
    for metric in (ef.METRIC_COSINE_SK, ef.METRIC_COSINE_G1, ef.METRIC_L2):
        results = []
        for R in (1, 3):
            scores = torch.empty((R, B), dtype=torch.float64, device="cuda")
            idxs = torch.empty((R, B), dtype=torch.int64, device="cuda")
            bounds = np.linspace(0, n, R + 1).astype(int)
            for r in range(R):
                lo, hi = int(bounds[r]), int(bounds[r + 1])
                gp = torch.empty((hi - lo, k), dtype=torch.float64, device="cuda")
                gn = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
                ef._lib.check(L.ef_gallery_prepare_device(g[lo:hi].data_ptr(), k, hi - lo, k, metric, gp.data_ptr(), k,
                                                          gn.data_ptr(), _stream(torch)), "prepare")
                work = torch.empty(L.ef_match_work_bytes(B, hi - lo) + 16, dtype=torch.uint8, device="cuda")
                ef._lib.check(L.ef_match_device(p.data_ptr(), k, B, k, gp.data_ptr(), k, gn.data_ptr(), hi - lo, lo, metric,
                                                scores[r].data_ptr(), idxs[r].data_ptr(), work.data_ptr(), _stream(torch)), "match")
            bs = torch.empty(B, dtype=torch.float64, device="cuda"); bi = torch.empty(B, dtype=torch.int64, device="cuda")
            ef._lib.check(L.ef_match_reduce_device(scores.data_ptr(), idxs.data_ptr(), R, B, metric, bs.data_ptr(),
                                                   bi.data_ptr(), _stream(torch)), "reduce")
            results.append((bs.cpu().numpy(), bi.cpu().numpy()))
        assert np.array_equal(results[0][1], results[1][1])
        np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-13)
        if metric == ef.METRIC_L2:
            d2 = ((P[:, None, :] - G[None]) ** 2).sum(-1)
            assert np.array_equal(results[0][1], d2.argmin(1))
        else:
            Pn = P / np.linalg.norm(P, axis=1, keepdims=True); Gn = G / np.linalg.norm(G, axis=1, keepdims=True)
            assert np.array_equal(results[0][1], (Pn @ Gn.T).argmax(1))
        assert results[0][1][-1] == 100                                   # three identical rows: lowest index wins


def _gram_tc(torch, L, x, N, D, side, G, d0=0, d1=None):
    d1 = D if d1 is None else d1
    wb = int(L.ef_gram_u8_tc_work_bytes(N, D, side))
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    st = L.ef_gram_u8_tc_device(x.data_ptr(), x.stride(0), N, D, d0, d1, side, G.data_ptr(), work.data_ptr(), wb,
                                _stream(torch))
    ef._lib.check(st, "ef_gram_u8_tc_device")
    torch.cuda.synchronize()
    assert int(work[:4].view(torch.int32).item()) == 0, "tcgen05 Gram pipeline timed out"


@pytest.mark.parametrize("N,D,side", [
    (229, 10000, 0),      # the shipped snapshot Gram (stream-K + RED)
    (150, 1000, 1),       # covariance side through the transposed copy
    (1, 16, 0), (5, 33, 1), (130, 257, 0), (300, 4096, 1),      # ragged tiles, K tails, n = 1
    (64, 70000, 0),       # K > 32768: s32 segments flushed into int64
    (40, 5000, 1),        # D x D = 5000 x 5000: whole-tile schedule (>= 2 tiles per SM), exclusive read-add-write
    (1000, 640, 1),       # MN-major operands straight from X: several K stages, ragged pixel blocks (640 = 2.5 x 256)
    (40000, 512, 1),      # MN-major, K = 40000 samples > 32768: two s32 segments per tile
    (700, 200, 1),        # D < 256: the transposed-copy fallback
])
def test_tensor_core_gram_is_exact(N, D, side):
    """tcgen05 kind::i8 SYRK == numpy int64 X X^T / X^T X, bit for bit, including the += contract."""
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(N * 7 + D)
    ld = (D + 15) // 16 * 16
    Xp = np.zeros((N, ld), dtype=np.uint8)
    Xp[:, :D] = rng.integers(0, 256, (N, D), dtype=np.uint8)
    if N >= 64 and side == 0:
        Xp[3, :D] = 255                                  # worst case magnitude on a full row
    x = torch.from_numpy(Xp).cuda()[:, :D]
    # reference through float64 BLAS: every sum is < 255^2 * 70 000 < 2^53, so the float64 product is the exact integer
    Xi = Xp[:, :D].astype(np.float64)
    n = N if side == 0 else D
    want = (Xi @ Xi.T if side == 0 else Xi.T @ Xi).astype(np.int64)
    G = torch.zeros((n, n), dtype=torch.int64, device="cuda")
    _gram_tc(torch, L, x, N, D, side, G)
    assert np.array_equal(G.cpu().numpy(), want)
    _gram_tc(torch, L, x, N, D, side, G)                 # accumulates
    assert np.array_equal(G.cpu().numpy(), 2 * want)
    if side == 0 and D >= 64:                             # pixel sub-ranges (column-sharded snapshot Gram)
        G.zero_()
        cut = (D // 2) // 16 * 16
        _gram_tc(torch, L, x, N, D, 0, G, 0, cut)
        _gram_tc(torch, L, x, N, D, 0, G, cut, D)
        assert np.array_equal(G.cpu().numpy(), want)


def test_tensor_core_gram_matches_cuda_core_gram_all_255():
    """Largest possible s32 segment sums (all pixels 255) agree between the tcgen05 and the dp4a kernels."""
    torch = require_gpu()
    L = ef._lib.lib()
    N, D = 140, 40000
    x = torch.full((N, D), 255, dtype=torch.uint8, device="cuda")
    G1 = torch.zeros((N, N), dtype=torch.int64, device="cuda")
    G2 = torch.zeros((N, N), dtype=torch.int64, device="cuda")
    _gram_tc(torch, L, x, N, D, 0, G1)
    ef._lib.check(L.ef_gram_u8_device(x.data_ptr(), D, N, D, 0, D, 0, G2.data_ptr(), _stream(torch)), "gram")
    assert torch.equal(G1, G2) and int(G1[0, 0].item()) == D * 255 * 255


@pytest.mark.parametrize("metric", [ef.METRIC_COSINE_SK, ef.METRIC_COSINE_G1, ef.METRIC_L2])
@pytest.mark.parametrize("n,k,B", [(30_000, 128, 300), (5000, 24, 77), (1, 7, 3), (257, 1, 129), (70_001, 50, 40)])
def test_tensor_core_matcher_equals_float64_scan(metric, n, k, B):
    """ef_match_tc_device (float16 tcgen05 filter + exact re-score) returns bit-identical (score, index) to the float64
    scan ef_match_device, including duplicates (lowest index), near-duplicates, zero rows and a zero query.  The L2
    metric (sum of squared differences, lowest wins) goes through the same GEMM with one extra component."""
    torch = require_gpu()
    L = ef._lib.lib()
    rng = np.random.default_rng(n + 31 * k + metric)
    G = rng.normal(size=(n, k)) / np.arange(1, k + 1)
    if n > 4200:
        G[4000] = G[100]; G[n - 1] = G[100]                    # exact duplicates
        G[200] = G[300] * (1 + 1e-9); G[201] = G[300] + 1e-7 * rng.normal(size=k)   # near duplicates
        G[400] = 0.0                                            # zero row
        G[401] = G[402] * (3.0 if metric == ef.METRIC_L2 else 1e6)   # (L2 keys are relative to the largest norm)
    P = G[rng.integers(0, n, B)] + 0.02 * rng.normal(size=(B, k)) / np.arange(1, k + 1)
    zero_query = B > 2 and n <= 60_000                          # (it ties with every row: n candidates of the 65 536 list)
    cosine = metric != ef.METRIC_L2
    if B > 2:
        if zero_query:
            P[1] = 0.0
        P[2] = G[min(100, n - 1)]
    g = torch.from_numpy(G).cuda(); p = torch.from_numpy(P).cuda()
    gp = torch.empty_like(g); gn = torch.empty(n, dtype=torch.float64, device="cuda")
    ef._lib.check(L.ef_gallery_prepare_device(g.data_ptr(), k, n, k, metric, gp.data_ptr(), k, gn.data_ptr(), _stream(torch)), "prep")
    # float64 scan
    s_ref = torch.empty(B, dtype=torch.float64, device="cuda"); i_ref = torch.empty(B, dtype=torch.int64, device="cuda")
    work = torch.empty(L.ef_match_work_bytes(B, n) + 16, dtype=torch.uint8, device="cuda")
    ef._lib.check(L.ef_match_device(p.data_ptr(), k, B, k, gp.data_ptr(), k, gn.data_ptr(), n, 1000, metric, s_ref.data_ptr(),
                                    i_ref.data_ptr(), work.data_ptr(), _stream(torch)), "match")
    # tensor-core filter + exact re-score
    img = torch.empty(L.ef_match_tc_image_bytes_metric(n, k, metric), dtype=torch.uint8, device="cuda")
    assert L.ef_match_tc_image_bytes(n, k) == L.ef_match_tc_image_bytes_metric(n, k, ef.METRIC_COSINE_G1)
    ef._lib.check(L.ef_match_tc_prepare_device(gp.data_ptr(), k, gn.data_ptr(), n, k, metric, img.data_ptr(), _stream(torch)), "tc prep")
    wb = L.ef_match_tc_work_bytes(B, n, k)
    wtc = torch.empty(wb, dtype=torch.uint8, device="cuda")
    s_tc = torch.empty(B, dtype=torch.float64, device="cuda"); i_tc = torch.empty(B, dtype=torch.int64, device="cuda")
    ef._lib.check(L.ef_match_tc_device(p.data_ptr(), k, B, k, gp.data_ptr(), k, gn.data_ptr(), img.data_ptr(), n, 1000, metric,
                                       s_tc.data_ptr(), i_tc.data_ptr(), wtc.data_ptr(), wb, _stream(torch)), "tc match")
    flags = (C.c_int32 * 3)()
    ef._lib.check(L.ef_match_tc_flags(wtc.data_ptr(), flags), "flags")
    assert flags[0] == 0 and flags[2] == 0, list(flags)
    assert flags[1] >= B - 1                                    # at least one survivor per (non-degenerate) query
    assert torch.equal(i_tc, i_ref)
    assert torch.equal(s_tc, s_ref)
    # the zero query ties with every row (all cosines are 0), everything else keeps a handful of survivors
    # (a zero query under L2 is not a tie: its distance to row j is |g_j|^2)
    if k > 1:                                                    # with k = 1 every cosine is +-1: half the gallery ties
        assert flags[1] < 50 * B + 2000 + (n if zero_query and cosine else 0), \
            f"filter too loose: {flags[1]} candidates for {B} queries"


def test_tensor_core_matcher_overflow_falls_back():
    """A gallery of identical rows puts every row inside the band: the candidate list overflows, the flag is raised and
    ShardedGallery answers from the float64 scan (lowest index wins)."""
    torch = require_gpu()
    n, k, B = 300_000, 16, 300
    G = np.tile(np.linspace(1.0, 2.0, k), (n, 1))
    P = np.tile(np.linspace(1.0, 2.0, k), (B, 1)) * np.linspace(0.5, 3.0, B)[:, None]
    sg = ef.dist.ShardedGallery(G, 7, ef.METRIC_COSINE_SK)
    score, index = sg.match_local(torch.from_numpy(P).cuda())
    assert sg.last_flags["overflow"] == 1
    assert torch.all(index == 7) and torch.allclose(score, torch.ones_like(score), atol=1e-12)


def test_tensor_core_matcher_l2_sharded_gallery_and_outlier_norm():
    """ShardedGallery with the L2 metric uses the tensor-core filter and answers exactly like the float64 scan; one gallery
    row with a norm 10^6 times the others squeezes every other key into the error band -- the list overflows and the
    float64 scan answers (same results, slower)."""
    torch = require_gpu()
    rng = np.random.default_rng(99)
    n, k, B = 200_000, 64, 500
    G = rng.normal(size=(n, k)) / np.sqrt(np.arange(1, k + 1))
    P = G[rng.integers(0, n, B)] + 0.05 * rng.normal(size=(B, k))
    P[3] *= 1e-3; P[4] *= 40.0                                   # queries far inside / outside the gallery's shell
    p = torch.from_numpy(P).cuda()
    ref = ef.dist.ShardedGallery(G, 11, ef.METRIC_L2, use_tensor_cores=False)
    s_ref, i_ref = ref.match_local(p)
    sg = ef.dist.ShardedGallery(G, 11, ef.METRIC_L2)
    assert sg.image is not None
    s, i = sg.match_local(p)
    assert sg.last_flags == {"timeout": 0, "candidates": sg.last_flags["candidates"], "overflow": 0}
    assert sg.last_flags["candidates"] < 20 * B
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref)
    d2 = ((P - G[i_ref.cpu().numpy() - 11]) ** 2).sum(-1)
    np.testing.assert_allclose(s.cpu().numpy(), d2, rtol=1e-12)
    G[777] *= 1e6
    sg2 = ef.dist.ShardedGallery(G, 11, ef.METRIC_L2)
    ref2 = ef.dist.ShardedGallery(G, 11, ef.METRIC_L2, use_tensor_cores=False)
    s2, i2 = sg2.match_local(p)
    s2r, i2r = ref2.match_local(p)
    assert sg2.last_flags["overflow"] == 1
    assert torch.equal(i2, i2r) and torch.equal(s2, s2r)
