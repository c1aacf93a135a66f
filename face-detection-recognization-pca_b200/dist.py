"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

Three natural shardings, each with at most one exchange step:
  * queries / frames      independent -> no collective (see bench.py --gpus N)
  * large gallery         contiguous row shards; per-shard top-1, one all-gather of (score, index), local reduce
  * training rows         per-rank exact integer column sums + integer Gram, one all-reduce(SUM) of int64

The collectives move exact integers or (score, index) pairs, so the result is bit identical to the single-GPU run for
any world size: integer sums are order independent and candidate reduction breaks ties towards the smallest GLOBAL
gallery row, exactly like np.argmax on the unsharded gallery.

Functions that only route tensors (shard_bounds, allgather_candidates, reduce_candidates, allreduce_exact) work on
CPU tensors with the gloo backend too; that is how the host logic is tested without GPUs.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import METRIC_COSINE_G1, METRIC_COSINE_SK, METRIC_L2, check


def shard_bounds(n, world, rank):
    """Contiguous partition of n rows: rank r owns [lo, hi); sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allgather_candidates(score, index, group=None):
    """[B] score (float64) and [B] global index (int64) of this rank -> ([R, B], [R, B]) on every rank."""
    import torch
    import torch.distributed as dist
    world, _ = _world(group)
    if world == 1:
        return score[None].clone(), index[None].clone()
    n = score.numel()
    scores = torch.empty(world * n, dtype=score.dtype, device=score.device)       # flat: gloo and nccl both accept it
    idxs = torch.empty(world * n, dtype=index.dtype, device=index.device)
    dist.all_gather_into_tensor(scores, score.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(idxs, index.contiguous().view(-1), group=group)
    return scores.view(world, n), idxs.view(world, n)


def reduce_candidates(scores, indices, metric):
    """Best candidate per query over R shards: higher cosine / lower L2 wins, ties -> smallest global index,
    shards that returned index -1 (empty) are ignored.  CUDA tensors use ef_match_reduce_device; CPU tensors the same
    rule in torch ops (host-logic tests)."""
    import torch
    R, B = scores.shape
    if scores.is_cuda:
        out_s = torch.empty(B, dtype=torch.float64, device=scores.device)
        out_i = torch.empty(B, dtype=torch.int64, device=scores.device)
        stream = C.c_void_p(torch.cuda.current_stream(scores.device).cuda_stream)
        check(_lib.lib().ef_match_reduce_device(scores.contiguous().data_ptr(), indices.contiguous().data_ptr(), R, B,
                                                metric, out_s.data_ptr(), out_i.data_ptr(), stream),
              "ef_match_reduce_device")
        return out_s, out_i
    worst = float("inf") if metric == METRIC_L2 else -float("inf")
    s = torch.where(indices < 0, torch.full_like(scores, worst), scores)
    best = s.min(dim=0).values if metric == METRIC_L2 else s.max(dim=0).values
    big = torch.iinfo(torch.int64).max
    cand = torch.where((s == best[None]) & (indices >= 0), indices, torch.full_like(indices, big))
    idx = cand.min(dim=0).values
    idx = torch.where(idx == big, torch.full_like(idx, -1), idx)
    return best, idx


def allreduce_exact(t, group=None):
    """SUM all-reduce of an integer tensor in place (exact, order independent)."""
    import torch.distributed as dist
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardedGallery:
    """This rank's contiguous slice of a large gallery, resident on the device.

    gallery_shard: [n_local, k] float64 (numpy or CUDA tensor); index_base: global row of its first row.
    match(features) returns the GLOBAL (score, index) per query, identical on every rank.
    """

    def __init__(self, gallery_shard, index_base, metric=METRIC_COSINE_SK, group=None, use_tensor_cores=True):
        import torch
        L = _lib.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        g = gallery_shard if torch.is_tensor(gallery_shard) else torch.from_numpy(
            np.ascontiguousarray(gallery_shard, dtype=np.float64))
        g = g.to(dev, dtype=torch.float64).contiguous()
        self.n, self.k = int(g.shape[0]), int(g.shape[1])
        self.index_base, self.metric, self.group = int(index_base), metric, group
        self.prepared = torch.empty_like(g)
        self.norms = torch.zeros(max(self.n, 1), dtype=torch.float64, device=dev)
        if self.n:
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(L.ef_gallery_prepare_device(g.data_ptr(), self.k, self.n, self.k, metric, self.prepared.data_ptr(),
                                              self.k, self.norms.data_ptr(), stream), "ef_gallery_prepare_device")
        self._L, self._work = L, None
        # tensor-core filter (k <= 128): float16 image of this shard, built once
        self.image = None
        self.last_flags = None
        if use_tensor_cores and self.n and self.k <= 128:
            self.image = torch.empty(int(L.ef_match_tc_image_bytes_metric(self.n, self.k, metric)), dtype=torch.uint8,
                                     device=dev)
            check(L.ef_match_tc_prepare_device(self.prepared.data_ptr(), self.k, self.norms.data_ptr(), self.n, self.k,
                                               metric, self.image.data_ptr(), stream), "ef_match_tc_prepare_device")
            self._work_tc = None

    def match_local(self, features):
        """Top-1 of the queries against this shard only: (score [B], global index [B]; -1 when the shard is empty)."""
        import torch
        p = features.contiguous()
        B = int(p.shape[0])
        score = torch.empty(B, dtype=torch.float64, device=p.device)
        index = torch.full((B,), -1, dtype=torch.int64, device=p.device)
        if self.n == 0 or B == 0:
            score.fill_(float("inf") if self.metric == METRIC_L2 else -float("inf"))
            return score, index
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        if self.image is not None:
            # filter on tensor cores + exact float64 re-score of the survivors (bit identical to the float64 scan)
            wb = int(self._L.ef_match_tc_work_bytes(B, self.n, self.k))
            if self._work_tc is None or self._work_tc.numel() < wb:
                self._work_tc = torch.empty(wb, dtype=torch.uint8, device=p.device)
            check(self._L.ef_match_tc_device(p.data_ptr(), p.stride(0), B, self.k, self.prepared.data_ptr(), self.k,
                                             self.norms.data_ptr(), self.image.data_ptr(), self.n, self.index_base,
                                             self.metric, score.data_ptr(), index.data_ptr(), self._work_tc.data_ptr(), wb,
                                             stream), "ef_match_tc_device")
            flags = (C.c_int32 * 3)()
            check(self._L.ef_match_tc_flags(self._work_tc.data_ptr(), flags), "ef_match_tc_flags")   # synchronises
            self.last_flags = {"timeout": flags[0], "candidates": flags[1], "overflow": flags[2]}
            if flags[0]:
                raise _lib.EigenfacesError(_lib.EF_ERR_CUDA, "ef_match_tc_device", "tcgen05 pipeline timed out")
            if not flags[2]:
                return score, index
            # degenerate gallery (candidate list overflow): the float64 scan below recomputes everything
        need = int(self._L.ef_match_work_bytes(B, self.n)) + 16
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.uint8, device=p.device)
        stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        check(self._L.ef_match_device(p.data_ptr(), p.stride(0), B, self.k, self.prepared.data_ptr(), self.k,
                                      self.norms.data_ptr(), self.n, self.index_base, self.metric, score.data_ptr(),
                                      index.data_ptr(), self._work.data_ptr(), stream), "ef_match_device")
        return score, index

    def match(self, features, timings=False):
        ph = _Phases(timings)
        ph.mark("start")
        score, index = self.match_local(features)
        ph.mark("local_top1")
        scores, idxs = allgather_candidates(score, index, self.group)        # B x 16 bytes per rank over NVLink
        ph.mark("allgather")
        out = reduce_candidates(scores, idxs, self.metric)
        ph.mark("reduce")
        if timings:
            self.last_timings = ph.result()
        return out



# ------------------------------------------------------------------------------ top-k eigenpairs of a large matrix
_TC_WORK = {}


def _dgemm(L, stream, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, Cm, ldc, split=False):
    """C = alpha A B + beta C (strided views).  FP64 tensor-core kernel (ef_dgemm_tc_device) unless EF_DGEMM_TC=0 or no
    stride of an operand is 1; split=True lets a small output with a long K use split-K (never for the row-sharded
    covariance products: their summation order must not depend on how the rows are distributed)."""
    import os
    if os.environ.get("EF_DGEMM_TC", "1") != "0" and (sak == 1 or sam == 1) and (sbk == 1 or sbn == 1):
        splits, work = 1, None
        tiles = -(-M // 128) * -(-N // 128)
        if split and tiles < 74 and K >= 1024:
            import torch
            splits = max(1, min(K // 256, -(-148 // tiles)))
            need = int(L.ef_dgemm_tc_work_bytes(M, N, splits))
            key = (Cm.device.index, )
            if key not in _TC_WORK or _TC_WORK[key].numel() < need:
                _TC_WORK[key] = torch.empty(max(need, 1 << 22), dtype=torch.uint8, device=Cm.device)
            work = _TC_WORK[key].data_ptr()
        check(L.ef_dgemm_tc_device(M, N, K, float(alpha), A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, float(beta),
                                   Cm.data_ptr(), ldc, splits, work, stream), "ef_dgemm_tc_device")
        return
    check(L.ef_dgemm_device(M, N, K, float(alpha), A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, float(beta),
                            Cm.data_ptr(), ldc, stream), "ef_dgemm_device")


def _jacobi(L, stream, H, work):
    """Symmetric m x m eigendecomposition on the device (destroys H): (evals descending [m], evecs [m, m] rows)."""
    import torch
    m = H.shape[0]
    evals = torch.empty(m, dtype=torch.float64, device=H.device)
    evecs = torch.empty((m, m), dtype=torch.float64, device=H.device)
    check(L.ef_eigh_jacobi_device(H.data_ptr(), m, evals.data_ptr(), evecs.data_ptr(), work.data_ptr(), 0, 0.0, None, None,
                                  stream), "ef_eigh_jacobi_device")
    return evals, evecs


def eigh_topk_device(Cm, k, tol=1e-11, max_outer=40, degree=None, block=None, seed=1234, group=None, orth_method=None,
                     lock=None):
    """Largest k eigenpairs of a symmetric positive semi-definite CUDA float64 matrix Cm [n, n] that is too large for the
    Jacobi solver (config 4: the 10 000 x 10 000 covariance, k = 256): Chebyshev-filtered subspace iteration with
    Rayleigh-Ritz and LOCKING (Zhou & Saad's scaled filter).  Replaces np.linalg.eigh(cov) + descending sort + top-k of
    useless/train.py:103-116 for large D.

    Every dense product is ef_dgemm_tc_device (FP64 tensor-core path), every small eigenproblem (block x block, block <= 320) the cluster-resident
    Jacobi kernel, the re-orthonormalisation after a filter application CholeskyQR2 (ef_chol_inverse_device: one CTA,
    ~0.1 ms; the Gram-matrix Jacobi route -- robust against rank loss, 10 ms per pass -- takes over when a pivot is not
    positive); torch only allocates and does O(n block) vector work.

    Locking: the leading run of Ritz pairs whose residual is below the tolerance leaves the block; the remaining
    columns iterate on the DEFLATED operator C - Ql diag(theta_l) Ql^T (two thin products per application), whose locked
    directions sit at eigenvalue ~0, inside the damped interval.  That is what makes a high filter degree usable: the
    filter's dynamic range is T_d(t_top) / T_d(t_wanted) with t = 2 lambda / beta - 1, and with the few large outliers of
    a face covariance (lambda_1 / lambda_256 = 60 on config 4) still in the block anything above degree 8 pushes the
    rounding-level components along them past the wanted ones (measured: degrees 12 ... 48 never converge).  Once the
    outliers are locked the top of the active block is within 10 % of beta and the degree is chosen from the range the
    float64 mantissa allows (`degree` = None: exp(30) of dynamic range, between 8 and 40; an int or a sequence -- last
    entry repeating -- fixes it per outer iteration).

    Stops when k pairs are locked (|C q - theta q| <= tol * theta_1 each), after max_outer outer iterations, or when the
    residual stagnates (info["stagnated"]: near-degenerate trailing eigenvalues, e.g. planted factors below the noise
    floor).  With a process group the dominant cost -- the n x n by n x block covariance products -- is SHARDED: rank r
    multiplies its contiguous block of rows of C and one all-gather (n x block float64 over NVLink) reassembles the
    product; every element is computed by the same kernel with the same summation order whoever owns its row, so all
    ranks hold bit-identical iterates (and the same result as a single GPU).  The block x block work stays replicated.
    Returns (evals [k] descending, evecs [n, k] orthonormal columns, info dict)."""
    import math
    import os

    import torch
    L = _lib.lib()
    dev = Cm.device
    n = int(Cm.shape[0])
    world, rank = _world(group)
    rows_per = -(-n // world)
    lo, hi = min(n, rank * rows_per), min(n, (rank + 1) * rows_per)
    k = int(k)
    m = int(block) if block else min(n, k + max(32, k // 4))      # k = 256 -> 320: the cluster Jacobi's largest size
    m = min(m, n)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    work = torch.empty(int(L.ef_eigh_work_bytes(m)), dtype=torch.uint8, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    method = orth_method or os.environ.get("EF_SUBSPACE_ORTH", "chol")
    if lock is None:
        lock = os.environ.get("EF_SUBSPACE_LOCK", "1") != "0"
    if degree is None and os.environ.get("EF_SUBSPACE_DEGREES"):
        degree = os.environ["EF_SUBSPACE_DEGREES"]
    if isinstance(degree, str):
        degree = [int(v) for v in degree.split(",")]
    degrees = None if degree is None else ([int(degree)] if isinstance(degree, int) else [int(v) for v in degree])
    if degrees is None and not lock:
        degrees = [8]
    info = {"outer": 0, "products": 0, "product_columns": 0, "block": m, "residual": None, "stagnated": False,
            "degrees": [], "locked_per_outer": []}
    gather_bufs = {}

    def cprod(alpha, Yin, beta, Yout):
        """Yout = alpha * C Yin + beta * Yout (any block width), rows of C sharded over the group."""
        w = int(Yin.shape[1])
        info["product_columns"] += w
        if world == 1:
            _dgemm(L, stream, n, w, n, alpha, Cm, n, 1, Yin, w, 1, beta, Yout, w)
            return
        import torch.distributed as dist
        if hi > lo:
            _dgemm(L, stream, hi - lo, w, n, alpha, Cm[lo:hi], n, 1, Yin, w, 1, beta, Yout[lo:hi], w)
        if w not in gather_bufs:
            gather_bufs[w] = torch.empty((world * rows_per, w), **f64)
        gather_buf = gather_bufs[w]
        mine = gather_buf[rank * rows_per:(rank + 1) * rows_per]
        mine.zero_()
        mine[:hi - lo].copy_(Yout[lo:hi])
        dist.all_gather_into_tensor(gather_buf.view(-1), mine.reshape(-1).clone(), group=group)
        Yout.copy_(gather_buf[:n])
        info["allgathers"] = info.get("allgathers", 0) + 1

    Ql = torch.empty((n, 0), **f64)          # locked Ritz vectors (columns) and values
    thl = torch.empty(0, **f64)

    def project_locked(Y, scale=None, alpha=-1.0):
        """Y += alpha * Ql diag(scale) Ql^T Y (scale None: the plain projector)."""
        nl, w = int(Ql.shape[1]), int(Y.shape[1])
        if nl == 0:
            return
        T = torch.empty((nl, w), **f64)
        _dgemm(L, stream, nl, w, n, 1.0, Ql, 1, nl, Y, w, 1, 0.0, T, w, split=True)  # Ql^T Y
        if scale is not None:
            T.mul_(scale[:, None])
        _dgemm(L, stream, n, w, nl, alpha, Ql, nl, 1, T, w, 1, 1.0, Y, w)           # Y += alpha Ql T

    def dprod(alpha, Yin, beta, Yout):
        """Product with the deflated operator C - Ql diag(thl) Ql^T."""
        cprod(alpha, Yin, beta, Yout)
        nl, w = int(Ql.shape[1]), int(Yin.shape[1])
        if nl:
            T = torch.empty((nl, w), **f64)
            _dgemm(L, stream, nl, w, n, 1.0, Ql, 1, nl, Yin, w, 1, 0.0, T, w, split=True)
            T.mul_(thl[:, None])
            _dgemm(L, stream, n, w, nl, -alpha, Ql, nl, 1, T, w, 1, 1.0, Yout, w)

    def orth(Y):
        """Orthonormal basis of span(Y): G = Y^T Y = W^T diag(g) W, Q = Y W^T diag(g^-1/2).  The columns are brought
        to unit norm first: after the filter they are nearly orthogonal Ritz directions whose NORMS span many orders of
        magnitude, and the Gram matrix of the unscaled block would lose the small ones."""
        w = int(Y.shape[1])
        norms = Y.norm(dim=0)
        Y = (Y / torch.where(norms > 0, norms, torch.ones_like(norms))[None, :]).contiguous()
        G = torch.empty((w, w), **f64)
        _dgemm(L, stream, w, w, n, 1.0, Y, 1, w, Y, w, 1, 0.0, G, w, split=True)    # Y^T Y
        g, W = _jacobi(L, stream, G, work)
        dead = g <= g[0] * 1e-28                                                    # directions lost to rounding
        scale = torch.where(dead, torch.zeros_like(g), g.clamp_min(1e-300).rsqrt())
        Ws = (W * scale[:, None]).contiguous()                                      # row i scaled by g_i^-1/2
        Q = torch.empty((n, w), **f64)
        _dgemm(L, stream, n, w, w, 1.0, Y, w, 1, Ws, 1, w, 0.0, Q, w)               # Y Ws^T
        n_dead = int(dead.sum())
        if n_dead:
            # rank loss: fresh random directions keep the block at full rank (zero columns would stay zero for ever);
            # the caller's second orthonormalisation pass makes them orthogonal to the rest
            Q[:, dead] = torch.randn((n, n_dead), generator=gen, **f64)
        return Q

    info_dev = torch.zeros(2, dtype=torch.int32, device=dev)

    def orth_chol(Y, slot):
        """One CholeskyQR pass: unit columns, G = Y^T Y = L L^T, Q = Y L^-T.  The pivot flag stays on the device."""
        w = int(Y.shape[1])
        norms = Y.norm(dim=0)
        Y = (Y / torch.where(norms > 0, norms, torch.ones_like(norms))[None, :]).contiguous()
        G = torch.empty((w, w), **f64)
        _dgemm(L, stream, w, w, n, 1.0, Y, 1, w, Y, w, 1, 0.0, G, w, split=True)    # Y^T Y
        Linv = torch.empty((w, w), **f64)
        check(L.ef_chol_inverse_device(G.data_ptr(), w, Linv.data_ptr(), info_dev[slot:].data_ptr(), stream),
              "ef_chol_inverse_device")
        Q = torch.empty((n, w), **f64)
        _dgemm(L, stream, n, w, w, 1.0, Y, w, 1, Linv, 1, w, 0.0, Q, w)             # Y L^-T
        return Q

    def orth2(Y):
        """Two passes: the Gram-matrix route squares the condition number, the second pass (CholQR2) restores
        orthonormality to rounding -- one pass leaves ~1e-9, which would floor the Ritz residual above the tolerance."""
        if method != "jacobi" and int(Y.shape[1]) <= 640:
            Q = orth_chol(orth_chol(Y, 0), 1)
            if not bool(info_dev.any()):                # one host read for both passes
                info["chol_orth"] = info.get("chol_orth", 0) + 1
                return Q
            info["chol_breakdowns"] = info.get("chol_breakdowns", 0) + 1
            info_dev.zero_()
        return orth(orth(Y))

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    history = []
    Q = orth2(torch.randn((n, m), generator=gen, **f64))
    lam = None
    theta1 = None
    for outer in range(max_outer):
        w = int(Q.shape[1])
        nl = int(Ql.shape[1])
        Y = torch.empty((n, w), **f64)
        dprod(1.0, Q, 0.0, Y)                                                       # Y = C' Q
        info["products"] += 1
        H = torch.empty((w, w), **f64)
        _dgemm(L, stream, w, w, n, 1.0, Q, 1, w, Y, w, 1, 0.0, H, w, split=True)    # H = Q^T C' Q
        H = ((H + H.T) * 0.5).contiguous()
        lam, W = _jacobi(L, stream, H, work)
        Qr = torch.empty((n, w), **f64)
        Yr = torch.empty((n, w), **f64)
        _dgemm(L, stream, n, w, w, 1.0, Q, w, 1, W, 1, w, 0.0, Qr, w)               # Ritz vectors Q W^T
        _dgemm(L, stream, n, w, w, 1.0, Y, w, 1, W, 1, w, 0.0, Yr, w)               # C' (Q W^T)
        if theta1 is None:
            theta1 = lam[0].clamp_min(1e-300).clone()
        want = k - nl                                                               # wanted pairs still in the block
        resid = (Yr[:, :want] - Qr[:, :want] * lam[None, :want]).norm(dim=0) / theta1
        res = float(resid.max())
        info.update(outer=outer + 1, residual=res)
        history.append(res)
        Q = Qr
        if res <= tol or w == n - nl:
            Ql = torch.cat([Ql, Qr[:, :want]], dim=1)
            thl = torch.cat([thl, lam[:want]])
            break
        n_lock = 0
        if lock:
            # leading run of converged pairs leaves the block
            ok = (resid <= tol).to(torch.int64)
            n_lock = int(torch.cumprod(ok, 0).sum())
            if n_lock:
                Ql = torch.cat([Ql, Qr[:, :n_lock]], dim=1).contiguous()
                thl = torch.cat([thl, lam[:n_lock]])
                Q = Qr[:, n_lock:].contiguous()
                Yr = Yr[:, n_lock:].contiguous()
                # the freshly locked directions leave C': C' q = C q - theta (q . q_l) q_l ~ C q for q orthogonal to them
                lam = lam[n_lock:]
                w -= n_lock
                history.clear()
        info["locked_per_outer"].append(n_lock)
        # stagnation: eigenvalues buried in a flat noise floor (gap << rounding of the products) cannot be resolved any
        # further; their invariant subspace is already captured to the reported residual
        if len(history) >= 10 and res > 0.9 * history[-7]:
            info["stagnated"] = True
            Ql = torch.cat([Ql, Q[:, :k - int(Ql.shape[1])]], dim=1)
            thl = torch.cat([thl, lam[:k - int(thl.shape[0])]])
            break
        # scaled Chebyshev filter: damps the unwanted interval [0, beta], beta = smallest Ritz value of the block
        beta, top = float(lam[w - 1]), float(lam[0])
        if not (top > beta > 0.0):
            beta = max(beta, 0.0) + 1e-3 * top
        if degrees is not None:
            deg = degrees[min(outer, len(degrees) - 1)]
        else:
            growth = math.acosh(max(2.0 * top / beta - 1.0, 1.0 + 1e-9))            # ln of the top's gain per degree
            deg = max(8, min(40, int(30.0 / growth)))
        info["degrees"].append(deg)
        e = c = 0.5 * beta
        sigma = e / (top - c)
        sigma1 = sigma
        Y1 = Yr.clone()
        Y1.sub_(Q * c).mul_(sigma1 / e)                                             # (C' Q - c Q) sigma1 / e
        Qp = Q
        for _ in range(2, deg + 1):
            sigma_new = 1.0 / (2.0 / sigma1 - sigma)
            Y2 = (Qp * (-sigma * sigma_new)).contiguous()
            Y2.sub_(Y1 * (2.0 * sigma_new * c / e))
            dprod(2.0 * sigma_new / e, Y1, 1.0, Y2)
            info["products"] += 1
            Qp, Y1, sigma = Y1, Y2, sigma_new
        project_locked(Y1)
        Q = orth2(Y1)
        project_locked(Q)                                                           # locked directions out to rounding
    else:
        Ql = torch.cat([Ql, Q[:, :k - int(Ql.shape[1])]], dim=1)
        thl = torch.cat([thl, lam[:k - int(thl.shape[0])]])
    info["locked"] = int(Ql.shape[1])
    order = torch.argsort(thl[:k], descending=True, stable=True)                    # locking order is by convergence
    return thl[:k][order].clone(), Ql[:, :k][:, order].contiguous(), info


class _Phases:
    """CUDA-event stopwatch for the phases of a multi-step device routine (events on torch's current stream; a NCCL
    collective issued through torch.distributed is ordered with that stream, so it is timed like a kernel)."""

    def __init__(self, enabled):
        self.enabled, self.marks = enabled, []

    def mark(self, name):
        if self.enabled:
            import torch
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def result(self):
        if not self.enabled or len(self.marks) < 2:
            return {}
        import torch
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b) * 1e-3
        out["total"] = self.marks[0][1].elapsed_time(self.marks[-1][1]) * 1e-3
        return out


def fit_gen1_sharded(X_local, n_total, n_components, group=None, solver="auto", timings=False):
    """Row-sharded manual_pca (useless/train.py:56-128) for N >= D (the covariance branch, e.g. 100 000 x 4096):
    every rank holds X_local uint8 [N_r, D] on its GPU.  Returns (eigenfaces [D,k], mean [D], projected_local [N_r,k],
    eigenvalues [k]) as CUDA float64 tensors; eigenfaces / mean / eigenvalues are identical on every rank.

    Per rank: exact integer column sums and Gram X_r^T X_r; ONE all-reduce(SUM) of int64 [D*D + D]; exact integer
    centring; replicated eigensolver (Jacobi for D <= 2048, Chebyshev-filtered subspace iteration for the top k above:
    `solver` = "auto" | "jacobi" | "subspace"); local projection of the local rows."""
    import torch
    L = _lib.lib()
    dev = X_local.device
    Nr, D = int(X_local.shape[0]), int(X_local.shape[1])
    if n_total < D:
        raise ValueError("fit_gen1_sharded covers the N >= D branch; small training sets fit on one GPU (fit_gen1)")
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ph = _Phases(timings)
    buf = torch.empty(D * D + D, dtype=torch.int64, device=dev)
    G, colsum = buf[:D * D], buf[D * D:]
    ph.mark("start")
    if Nr:
        check(L.ef_colsum_u8_device(X_local.data_ptr(), X_local.stride(0), Nr, D, colsum.data_ptr(), stream), "colsum")
        # exact integer X_r^T X_r on tensor cores (tcgen05 kind::i8), stored (no zero fill, no read-add-write of the
        # 8 D^2-byte result); the dp4a kernel covers unaligned buffers
        wb = int(L.ef_gram_u8_tc_work_bytes(Nr, D, 1))
        gwork = torch.empty(wb, dtype=torch.uint8, device=dev)
        st_g = L.ef_gram_u8_tc_store_device(X_local.data_ptr(), X_local.stride(0), Nr, D, 0, D, 1, G.data_ptr(),
                                            gwork.data_ptr(), wb, stream)
        if st_g == _lib.EF_ERR_UNSUPPORTED:
            G.zero_()
            st_g = L.ef_gram_u8_device(X_local.data_ptr(), X_local.stride(0), Nr, D, 0, D, 1, G.data_ptr(), stream)
        check(st_g, "gram")
    else:
        buf.zero_()
    ph.mark("gram")
    allreduce_exact(buf, group)
    ph.mark("allreduce")
    cov = torch.empty((D, D), dtype=torch.float64, device=dev)
    check(L.ef_gram_center_device(G.data_ptr(), D, 1, colsum.data_ptr(), int(n_total), 1.0 / (n_total - 1),
                                  cov.data_ptr(), None, stream), "center")
    ph.mark("center")
    k = min(int(n_components), D)
    if solver == "subspace" or (solver == "auto" and D > 2048):
        # large D (config 4: 10 000 pixels): only the top k eigenpairs, by filtered subspace iteration
        evals, E_top, solver_info = eigh_topk_device(cov, k, group=group)
        fit_gen1_sharded.last_solver_info = solver_info
        evecs = None
    else:
        evals = torch.empty(D, dtype=torch.float64, device=dev)
        evecs = torch.empty((D, D), dtype=torch.float64, device=dev)
        work = torch.empty(int(L.ef_eigh_work_bytes(D)), dtype=torch.uint8, device=dev)
        check(L.ef_eigh_jacobi_device(cov.data_ptr(), D, evals.data_ptr(), evecs.data_ptr(), work.data_ptr(), 0, 0.0, None,
                                      None, stream), "jacobi")
    ph.mark("solver")
    # tensor divisor: torch turns division by a Python scalar into a multiplication by the reciprocal on CUDA, which is
    # not the correctly rounded quotient np.mean (and the single-GPU fit) returns
    mean = colsum.to(torch.float64) / torch.full((1,), float(n_total), dtype=torch.float64, device=dev)
    E = evecs[:k].T.contiguous() if evecs is not None else E_top   # [D, k]
    Z = torch.empty((max(Nr, 1), D), dtype=torch.float64, device=dev)
    proj = torch.empty((Nr, k), dtype=torch.float64, device=dev)
    if Nr:
        check(L.ef_standardize_u8_device(X_local.data_ptr(), X_local.stride(0), Nr, D, mean.data_ptr(), None, None,
                                         Z.data_ptr(), D, stream), "center rows")
        _dgemm(L, stream, Nr, k, D, 1.0, Z, D, 1, E, k, 1, 0.0, proj, k)
    ph.mark("projection")
    fit_gen1_sharded.last_timings = ph.result()
    return E, mean, proj, evals[:k].clone()
