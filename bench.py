#!/usr/bin/env python
"""Benchmark of the Eigenfaces hot path (BASELINE.json metric: face crops/sec projected+matched; PCA fit seconds).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Workload C2 (BASELINE.json configs[1], the configuration the metric is quoted on; fits one GPU): a batch of 4096
synthetic 100x100 gray crops (D = 10 000) projected onto k = 10 eigenfaces and matched (cosine top-1 + threshold +
reconstruction error) against a 1024-row gallery.  One step = one batch.  At N > 1 every rank runs its own batches
(queries are data parallel, no collective on the data path): weak scaling, value = all crops / max-over-ranks time.

Prints ONE JSON line on rank 0 (see the driver contract in the task statement):
  value     device-resident throughput (inputs already in HBM, CUDA events on the launching stream)
  e2e       the same metric through the host-buffer C-ABI call (pinned host crops in, labels out, copies inside)
  roofline  dominant kernel (the projection) against the measured HBM peak
  cpu_baseline  the oracle port of the reference path timed on this box's host cores (rank 0, N = 1)
--impl reference times only the CPU port (the reference is pure Python + numpy/sklearn and cannot be installed
on the GPU box; the oracle restates it -- see DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, SIDE, K_COMP, N_GALLERY, N_BATCHES = 4096, 100, 10, 1024, 8
D = SIDE * SIDE
THRESHOLD = 0.8
ALGO_BYTES_PER_CROP = D + 4 * K_COMP + 8
METRIC = "face crops/sec projected+matched"
WORKLOAD = ("C2: B=4096 synthetic face-like 100x100 gray crops (D=10000) -> k=10 eigenfaces -> cosine top-1 vs "
            "1024-row gallery + threshold + reconstruction error")


def training_matrix():
    """The 229 x 10000 light training crops (golden fixture, travels with the repo); synthetic fallback."""
    path = os.path.join(ROOT, "tests", "golden", "gen1_light.npz")
    if os.path.exists(path):
        return np.load(path)["X_u8"], "eigenfaces fitted on tests/golden/gen1_light.npz (229 crops)"
    rng = np.random.default_rng(7)
    base = rng.normal(0, 1, (229, 24)) @ rng.normal(0, 1, (24, D))
    return np.clip(np.rint(128 + 18 * base + rng.normal(0, 6, (229, D))), 0, 255).astype(np.uint8), "synthetic basis"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "window": "timed regions + post-roll of the same step (>= 1.5 s under load)"}


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_port_setup(seed=20250820):
    """Model + one batch, everything on the CPU with the oracle (the reference's arithmetic)."""
    from oracle import gen1
    X, note = training_matrix()
    E, mu, _, lam = gen1.manual_pca(X.astype(np.float64), K_COMP)
    rng = np.random.default_rng(seed)
    G = gen1.project_batch(face_like_np(rng, E, mu, lam, N_GALLERY), E, mu)
    model = dict(eigenfaces=E, mean_face=mu, projected_data=G, person_name="bench")
    return model, lam, note


def face_like_np(rng, E, mu, lam, n):
    c = rng.normal(0, 1, (n, len(lam))) * np.sqrt(lam)
    return np.clip(np.rint(mu + c @ E.T + rng.normal(0, 8, (n, E.shape[0]))), 0, 255).astype(np.uint8)


def cpu_port_time(model, Q, budget_s, max_steps=None):
    """Batched numpy port of useless/scan.py:recognize_face over whole batches until the budget is spent."""
    from oracle import gen1
    gen1.recognize_batch(Q[:256], model, THRESHOLD)                  # warm-up (BLAS threads, page faults)
    times = []
    t_end = time.perf_counter() + budget_s
    while (time.perf_counter() < t_end and (max_steps is None or len(times) < max_steps)) or not times:
        t0 = time.perf_counter()
        gen1.recognize_batch(Q, model, THRESHOLD)
        times.append(time.perf_counter() - t0)
    return times


def cpu_literal_time(model, Q, n=32):
    """The reference's literal per-crop loop (Python loop over the gallery, useless/scan.py:100-132)."""
    from oracle import gen1
    t0 = time.perf_counter()
    for row in Q[:n]:
        gen1.recognize_face(row.flatten().astype(np.float64), model, THRESHOLD)
    return n / (time.perf_counter() - t0)


def run_reference(args, rank):
    if rank != 0:
        return
    model, lam, note = cpu_port_setup()
    rng = np.random.default_rng(1)
    Q = face_like_np(rng, model["eigenfaces"], model["mean_face"], lam, B)
    for _ in range(max(args.warmup, 1) - 1):
        cpu_port_time(model, Q, 0.0, 1)
    # bounded: at most --steps batches and at most ~120 s
    times = cpu_port_time(model, Q, 120.0, args.steps)
    crops_s = B * len(times) / sum(times)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": crops_s, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "basis": note},
        "cpu_baseline": {"value": crops_s, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full batches of {B} crops through the batched numpy port of "
                                   "useless/scan.py:recognize_face (oracle/gen1.py:recognize_batch), BLAS threads = all cores",
                         "literal_per_crop_crops_s": cpu_literal_time(model, Q)},
        "e2e": {"value": crops_s, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------ secondary measurements
def fit_and_gram_section(ef, torch, dev, peaks):
    """PCA fit seconds (the second half of BASELINE.json's metric) and the tensor-core Gram behind it.  Bounded: a few
    seconds.  fit: the shipped Gen-1 configuration (229 light crops x 10 000 pixels, k = 50) through ef_fit_gen1_host
    (H2D + Gram + Jacobi + back-projection + D2H), the oracle's manual_pca on the host cores beside it.  gram: the
    exact u8 x u8 tcgen05 SYRK at the config-4 per-GPU shape (12 500 rows x 10 000 pixels -> 10 000 x 10 000)."""
    import ctypes as C
    from oracle import gen1
    out = {}
    X, note = training_matrix()
    t0 = time.perf_counter(); ef.fit_gen1(X, 50); torch.cuda.synchronize()
    walls, gpu = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        info = ef.fit_gen1(X, 50)[4]
        walls.append(time.perf_counter() - t0); gpu.append(info["gpu_ms"])
    cpu = []
    Xf = X.astype(np.float64)
    for _ in range(3):
        t0 = time.perf_counter(); gen1.manual_pca(Xf, 50); cpu.append(time.perf_counter() - t0)
    out["fit"] = {"what": f"manual_pca {X.shape[0]}x{X.shape[1]} k=50 ({note})", "unit": "s",
                  "gpu_wall_s": min(walls), "gpu_device_s": min(gpu) * 1e-3, "jacobi_sweeps": info["sweeps"],
                  "cpu_port_s": min(cpu), "cpu_cores": os.cpu_count(),
                  "note": "gpu_wall_s = ef_fit_gen1_host call incl. H2D/D2H and allocations; cpu_port_s = oracle/gen1.py:manual_pca (numpy, all cores)"}
    L = ef._lib.lib()
    N, Dg = 12500, 10000
    x = torch.randint(0, 256, (N, Dg), dtype=torch.uint8, device=dev)
    G = torch.zeros((Dg, Dg), dtype=torch.int64, device=dev)
    wb = int(L.ef_gram_u8_tc_work_bytes(N, Dg, 1))
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run():
        ef._lib.check(L.ef_gram_u8_tc_device(x.data_ptr(), x.stride(0), N, Dg, 0, Dg, 1, G.data_ptr(), work.data_ptr(),
                                             wb, st), "ef_gram_u8_tc_device")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tiles = sum(min((256 * tj + 255) // 128 + 1, (Dg + 127) // 128) for tj in range((Dg + 255) // 256))
    executed = 2.0 * tiles * 128 * 256 * ((N + 127) // 128 * 128)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    out["gram"] = {"what": f"ef_gram_u8_tc_device side 1: X^T X of u8[{N},{Dg}] -> int64[{Dg},{Dg}] (transpose + tcgen05 "
                           "kind::i8 SYRK upper triangle + mirror), exact",
                   "ms": ms, "algorithmic_tops": 2.0 * Dg * Dg * N / ms / 1e9 / 2.0,
                   "executed_tops": executed / ms / 1e9,
                   "roofline": {"bound": "tensor", "achieved": executed / ms / 1e9, "peak": bf16_peak, "unit": "TOP/s (int8 ops) vs measured bf16 TFLOP/s",
                                "frac": executed / ms / 1e9 / bf16_peak,
                                "note": "algorithmic = N*D^2 (symmetric half); executed = upper-triangle tiles incl. the diagonal overlap; int8 nominal peak is 2x bf16"},
                   "flag": int(work[:4].view(torch.int32).item())}
    del x, G, work
    # K1: resize-active preprocess, ROI 100..300 px squares inside 1080p gray frames -> 100x100
    rng = np.random.default_rng(5150)
    F, H, W, nb = 8, 1080, 1920, 4096
    frames = torch.randint(0, 256, (F, H, W), dtype=torch.uint8, device=dev)
    side = rng.integers(100, 301, nb)
    bx = np.stack([rng.integers(0, F, nb), (rng.random(nb) * (W - side)).astype(np.int64),
                   (rng.random(nb) * (H - side)).astype(np.int64), side, side], axis=1).astype(np.int32)
    boxes = torch.from_numpy(bx).to(dev)
    outp = ef.engine.preprocess_device(frames, boxes, 100)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ef.engine.preprocess_device(frames, boxes, 100, out=outp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    bytes_alg = float((side.astype(np.int64) ** 2).sum() + nb * 10000)
    hbm = peaks.get("hbm_gbs", 6650.0)
    out["preprocess"] = {"what": "ef_preprocess: 4096 square ROIs (100..300 px) of 1080p gray frames -> 100x100 (bit exact cv2.resize INTER_LINEAR)",
                         "ms": ms, "crops_per_s": nb / ms * 1e3,
                         "roofline": {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                      "frac": bytes_alg / ms / 1e6 / hbm,
                                      "note": "algorithmic bytes = ROI pixels read once + 10 000 B written per crop"}}
    del frames, boxes, outp
    # config 3 on one GPU: 1 M gallery identities x k = 128, 4096 queries, tensor-core filter + exact float64 re-score
    gen = torch.Generator(device=dev); gen.manual_seed(1_000_003)
    n3, k3 = 1_000_000, 128
    lam3 = torch.arange(1, k3 + 1, device=dev, dtype=torch.float64) ** -2.0
    G3 = torch.randn((n3, k3), generator=gen, device=dev, dtype=torch.float64) * lam3.sqrt()
    truth = torch.randint(0, n3, (4096,), generator=gen, device=dev)
    P3 = G3[truth] + 0.05 * torch.randn((4096, k3), generator=gen, device=dev, dtype=torch.float64) * lam3.sqrt()
    sg = ef.dist.ShardedGallery(G3, 0, ef.METRIC_COSINE_SK)
    for _ in range(2):
        sg.match_local(P3)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        s3, i3 = sg.match_local(P3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    sg64 = ef.dist.ShardedGallery(G3, 0, ef.METRIC_COSINE_SK, use_tensor_cores=False)
    sg64.match_local(P3[:64]); torch.cuda.synchronize()
    e0.record(); s64, i64 = sg64.match_local(P3[:64]); e1.record(); torch.cuda.synchronize()
    ms64 = e0.elapsed_time(e1) * 4096 / 64
    f16_flops = 2.0 * 2 * 4096 * n3 * 384
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    out["large_gallery"] = {
        "what": "config 3 on one GPU: 4096 queries x 1 000 000 gallery rows x k = 128, cosine top-1 (ef_match_tc_device: "
                "tcgen05 f16 hi/lo filter GEMM, two passes, + exact float64 re-score of the survivors)",
        "ms_per_batch": ms, "queries_per_s": 4096 / ms * 1e3, "candidates_rescored": sg.last_flags["candidates"],
        "top1_accuracy_vs_planted": float((i3 == truth).double().mean()),
        "bit_identical_to_float64_scan_on_64_queries": bool(torch.equal(i64, i3[:64]) and torch.equal(s64, s3[:64])),
        "float64_scan_ms_per_batch_extrapolated": ms64,
        "roofline": {"bound": "tensor", "achieved": f16_flops / ms / 1e9, "peak": bf16_peak, "unit": "TFLOP/s (f16 filter GEMM, both passes)",
                     "frac": f16_flops / ms / 1e9 / bf16_peak}}
    del G3, P3, sg, sg64
    # config 4 shape on one GPU (N reduced to 25 000 rows to bound the run): tensor-core Gram, exact centring,
    # Chebyshev-filtered subspace iteration for the top 256 eigenpairs of the 10 000 x 10 000 covariance, projection
    N4, D4, R4, K4 = 25_000, 10_000, 300, 256
    gen.manual_seed(4242)
    F4 = torch.linalg.qr(torch.randn((D4, R4), generator=gen, device=dev, dtype=torch.float32))[0]
    sig4 = 40.0 * torch.arange(1, R4 + 1, device=dev, dtype=torch.float32) ** -0.7
    X4 = torch.empty((N4, D4), dtype=torch.uint8, device=dev)
    for i in range(0, N4, 5000):
        L4 = torch.randn((5000, R4), generator=gen, device=dev) * sig4
        X4[i:i + 5000] = (128 + L4 @ F4.T + 4.0 * torch.randn((5000, D4), generator=gen, device=dev)).round_().clamp_(0, 255).to(torch.uint8)
    ef.dist.fit_gen1_sharded(X4, N4, K4); torch.cuda.synchronize()
    t0 = time.perf_counter()
    E4, _, _, ev4 = ef.dist.fit_gen1_sharded(X4, N4, K4)
    torch.cuda.synchronize()
    info4 = dict(ef.dist.fit_gen1_sharded.last_solver_info)
    out["fit_large"] = {"what": f"config 4 shape, one GPU: manual_pca covariance branch on u8[{N4},{D4}], k = {K4} "
                                "(tensor-core Gram + integer centring + filtered subspace iteration + projection)",
                        "seconds": time.perf_counter() - t0, "solver": info4,
                        "orthonormality_error": float((E4.T @ E4 - torch.eye(K4, device=dev, dtype=torch.float64)).abs().max())}
    # ---- template-matching detector: one 640x480 camera frame against 4 persons x 5 template crops
    # x 3 scales = 60 TM_CCOEFF_NORMED maps + arg-max, the per-frame work of scan-template-v4.py:129-197
    rng = np.random.default_rng(640480)
    frame = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    tmpls = [rng.integers(0, 256, (int(rng.integers(80, 121)), int(rng.integers(80, 121))), dtype=np.uint8) for _ in range(20)]
    tmpls[3] = frame[200:300, 250:340].copy()
    matcher = ef.template.TemplateMatcher(tmpls)
    frame_dev = torch.from_numpy(frame).to(dev)
    res = matcher.match(frame_dev)
    torch.cuda.synchronize()
    l0 = ef.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        res = matcher.match(frame_dev)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    macs = sum((640 - w + 1) * (480 - h + 1) * w * h for (_, _, w, h) in matcher.jobs)
    tm = {"what": f"cv2.matchTemplate(TM_CCOEFF_NORMED) + minMaxLoc, 640x480 gray frame, {len(matcher.jobs)} (template, scale) jobs "
                  "of 64..144 px (scan-template-v4.py:129-197 per frame)",
          "ms_per_frame_device": e0.elapsed_time(e1) / 5, "ms_per_frame_wall": wall * 1e3, "launches_per_frame": (ef.launch_count() - l0) / 5,
          "exact_integer_tmac_per_s": macs / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12,
          "found_pasted_template_at": [res[10]["x"], res[10]["y"]] if res[10] else None}
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
        t0 = time.perf_counter()
        n_cpu = 0
        for ti, scale, w, h in matcher.jobs[:12]:                       # bounded sample: 12 of the 60 jobs
            st = cv2.resize(tmpls[ti], (w, h))
            r = cv2.matchTemplate(frame, st, cv2.TM_CCOEFF_NORMED)
            cv2.minMaxLoc(r)
            n_cpu += 1
        tm["cpu_cv2_ms_per_frame"] = (time.perf_counter() - t0) / n_cpu * len(matcher.jobs) * 1e3
        tm["cpu_note"] = f"cv2 {cv2.__version__} on {os.cpu_count()} host cores, {n_cpu} of the jobs timed and scaled to all"
    except ImportError:
        tm["cpu_cv2_ms_per_frame"] = None
    out["template_match"] = tm
    return out


# ------------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import eigenfaces_b200 as ef

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- model: fit the k=10 eigenfaces with the engine's own PCA fit (outside every timed region)
    X, note = training_matrix()
    E, mu, _, lam, fit_info = ef.fit_gen1(X, K_COMP)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20250820 + rank)
    E_t = torch.from_numpy(np.ascontiguousarray(E)).to(dev)
    mu_t = torch.from_numpy(mu).to(dev)
    sq_t = torch.from_numpy(np.sqrt(lam)).to(dev)

    def face_like(n):
        c = torch.randn((n, K_COMP), generator=gen, device=dev, dtype=torch.float64) * sq_t
        x = mu_t + c @ E_t.T
        x += 8.0 * torch.randn((n, D), generator=gen, device=dev, dtype=torch.float64)
        return x.round_().clamp_(0, 255).to(torch.uint8)

    ld = (D + 127) // 128 * 128
    proj_only = ef.Recognizer(E, mu, np.zeros((1, K_COMP)), metric=ef.METRIC_COSINE_G1, with_residual=False)
    gal_crops = torch.zeros((N_GALLERY, ld), dtype=torch.uint8, device=dev)
    gal_crops[:, :D] = face_like(N_GALLERY)
    G = proj_only.recognize_device(gal_crops, 0.0)["features"].cpu().numpy()
    proj_only.close()
    labels = (np.arange(N_GALLERY) % 4).astype(np.int32)
    rec = ef.Recognizer(E, mu, G, labels=labels, metric=ef.METRIC_COSINE_G1, with_residual=True)
    rec.reserve(B)

    batches = []
    for _ in range(N_BATCHES):
        xb = torch.zeros((B, ld), dtype=torch.uint8, device=dev)
        xb[:, :D] = face_like(B)
        batches.append(xb)
    out = rec.recognize_device(batches[0], THRESHOLD)       # allocates the output tensors once
    host_batches = []
    for i in range(2):
        hb = torch.empty((B, D), dtype=torch.uint8, pin_memory=True)
        hb.copy_(batches[i][:, :D])
        host_batches.append(hb.numpy())
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    def step(i):
        rec.recognize_device(batches[i % N_BATCHES], THRESHOLD, out=out)

    def step_pipelined(i):
        # serving loop: the launch of batch i also matches batch i-1 (ef_model_submit_device)
        rec.submit_device(batches[i % N_BATCHES], THRESHOLD, out=out)

    # ---- device-resident throughput, one call = one complete batch (results of batch i ready after call i)
    for i in range(args.warmup):
        step(i)
    barrier()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_unpipelined = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    # ---- device-resident throughput of the serving loop (the headline `value`): K submits + the final flush, all inside
    # the timed region; every batch is streamed, projected and matched completely, only the match of batch i runs in the
    # launch of batch i+1
    for i in range(args.warmup):
        step_pipelined(i)
    rec.flush_device()
    barrier()
    launches0 = ef.launch_count()
    e0.record()
    for i in range(args.steps):
        step_pipelined(i)
    rec.flush_device()
    e1.record()
    torch.cuda.synchronize()
    launches = ef.launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- dominant kernel (projection) timed with CUDA events on the launching stream, same steps
    rec.kernel_timing(True)
    for i in range(min(args.steps, 512)):
        step(i)
    torch.cuda.synchronize()
    n_calls, proj_ms, used_tc = rec.kernel_timing_read()
    rec.kernel_timing(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host crops in, results out, copies inside)
    for i in range(max(args.warmup, 3)):
        rec.recognize(host_batches[i % 2], THRESHOLD, want_features=False)
    barrier()
    e2e_steps = max(1, min(args.steps, 200)) if not args.profile else 1
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r = rec.recognize(host_batches[i % 2], THRESHOLD, want_features=False)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_sync_value = world * B * e2e_steps / e2e_s
    e2e_sync_ms = 1e3 * e2e_s / e2e_steps
    d2h = B * (8 + 4 + 4 + 8)
    # the serving form of the same call: ef_model_submit_host / ef_model_wait_host, two batches in flight -- every step
    # still uploads its 4096 crops from pinned host memory and reads its scores / indices / labels / residuals back
    # inside the timed region; the upload of step i+1 overlaps the kernels and the result copy of step i
    tk = [rec.submit(host_batches[0], THRESHOLD, want_features=False), None]
    for i in range(max(args.warmup, 3)):
        tk[(i + 1) % 2] = rec.submit(host_batches[(i + 1) % 2], THRESHOLD, want_features=False)
        rec.wait(tk[i % 2])
    rec.wait(tk[max(args.warmup, 3) % 2])
    barrier()
    t0 = time.perf_counter()
    tk = [rec.submit(host_batches[0], THRESHOLD, want_features=False), None]
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            tk[(i + 1) % 2] = rec.submit(host_batches[(i + 1) % 2], THRESHOLD, want_features=False)
        r = rec.wait(tk[i % 2])
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * e2e_steps / e2e_s

    # ---- post-roll under the same load so that nvidia-smi gets samples even when the timed region is short
    t_roll = time.perf_counter()
    i = 0
    while time.perf_counter() - t_roll < (0.0 if args.profile else 1.5):
        step(i); i += 1
        if i % 64 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    t_end = time.perf_counter()

    sanity = rec.recognize(host_batches[0][:64], THRESHOLD)
    assert np.isfinite(sanity.score).all() and (sanity.index >= 0).all()

    if rank == 0:
        clocks = sampler.stop(t_begin, t_end)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # one launch per step (single cluster kernel): the average launch duration over the timed region IS the step
        # time, launch gaps included (consecutive launches overlap through programmatic dependent launch, so bracketing
        # each launch with its own events would serialise them: that figure is kept as kernel_ms_isolated)
        one_launch = launches == args.steps + 1           # K pipelined submits + 1 flush launch
        kernel_ms = ms / launches if one_launch else proj_ms
        achieved = ALGO_BYTES_PER_CROP * B / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        extra = {}
        if world == 1 and not args.profile and not args.no_extras:
            extra = fit_and_gram_section(ef, torch, dev, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not args.profile:
            model, lam_c, _ = cpu_port_setup()
            Q = host_batches[0]
            times = cpu_port_time(model, Q, 12.0)
            cpu = {"value": B * len(times) / sum(times), "unit": "crops/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{len(times)} batches of {B} crops (same workload) through oracle/gen1.py:recognize_batch "
                             "(batched numpy port of useless/scan.py:recognize_face), all host cores via BLAS",
                   "literal_per_crop_crops_s": cpu_literal_time(model, Q)}
        line = {
            "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 x s8 digit planes -> s32 (exact), f64 combine + match",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "basis": note, "n_slices": 8, "threshold": THRESHOLD,
                       "l2": f"{N_BATCHES} distinct resident batches rotated ({N_BATCHES * B * ld / 1e6:.0f} MB > 126 MB L2)",
                       "call": "ef_model_submit_device per step + ef_model_flush_device at the end (inside the timed region)",
                       "projection_kernel": {3: "recognize_pipe_kernel: stream half (TMA + tcgen05 kind::i8 + DSMEM push + f64 features) of batch i and match half (tcgen05 f16 filter + exact f64 re-score) of batch i-1 in one launch, PDL",
                                             2: "recognize_cluster_kernel: TMA + tcgen05 kind::i8 + DSMEM push + tcgen05 f16 filter + exact f64 re-score (1 launch/step, PDL)",
                                             1: "project_tc_kernel (tcgen05 kind::i8, stream-K) + fused_epilogue_kernel",
                                             0: "project_dp4a_kernel (CUDA cores) + fused_epilogue_kernel"}[3 if launches == args.steps + 1 else int(used_tc)]},
            "unpipelined": {"value": world * B * args.steps / (ms_unpipelined * 1e-3), "ms_per_step": ms_unpipelined / args.steps,
                            "what": "same steps through ef_model_recognize_device (all results of a batch ready after its own call)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (ncu --set full,
                         # profiles/r1_h_summary.md): the crops are read once, nothing is re-read
                         "traffic": 43332864 if launches == args.steps + 1 else None,
                         "traffic_source": "profiles/r1_h_summary.md (ncu --set full, recognize_pipe_kernel<1,12>, bytes per launch)",
                         "kernel": ("recognize_pipe_kernel (whole step: stream + project batch i, match batch i-1)" if one_launch
                                    else "projection (digit-plane integer GEMM)"), "kernel_ms": kernel_ms,
                         "kernel_ms_isolated": proj_ms, "kernel_calls_timed": args.steps if one_launch else n_calls,
                         "how": ("device-timed region / launches (1 launch per step, CUDA events on the launching stream)"
                                 if one_launch else "CUDA event pair around every projection launch"),
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CROP * B, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": B * D, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "ef_model_submit_host / ef_model_wait_host (Recognizer.submit / wait), two batches in flight",
                    "synchronous": {"value": e2e_sync_value, "ms_per_step": e2e_sync_ms,
                                    "api": "ef_model_recognize_host (Recognizer.recognize), one call per step"},
                    "timer": "host perf_counter around the whole loop of calls, max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "fit_setup": {"what": "ef_fit_gen1_host 229x10000 k=10 (model setup, untimed)", "gpu_ms": fit_info["gpu_ms"],
                          "jacobi_sweeps": fit_info["sweeps"]},
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fit / Gram / preprocess side measurements")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: no CPU baseline, no clock post-roll, one e2e step (numbers are not bench values)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
