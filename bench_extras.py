"""Secondary measurements printed inside bench.py's JSON line (never the headline `value`): the other BASELINE.json
configs (C1 train-v5 single person, C3 large gallery sharded, C4 fit at scale sharded, C5 video frames), the PCA-fit
seconds of the metric, the per-kernel rooflines of K1 / K3, the shipped model shapes and the B = 1 drop-in latency.

Every section is bounded (seconds), times the device with CUDA events and, where BASELINE.md section 3 prescribes one,
times the CPU restatement (oracle/) beside it on the host cores.  Sections that take a process group run on EVERY rank
(they issue NCCL collectives); the others run on rank 0 of a single-GPU launch only.
"""
import ctypes as C
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_loop(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _max_over_ranks(torch, dist, world, dev, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def training_matrix(D=10000):
    """The 229 x 10000 light training crops (golden fixture, travels with the repo); synthetic fallback."""
    path = os.path.join(GOLDEN, "gen1_light.npz")
    if os.path.exists(path):
        return np.load(path)["X_u8"], "eigenfaces fitted on tests/golden/gen1_light.npz (229 crops)"
    rng = np.random.default_rng(7)
    base = rng.normal(0, 1, (229, 24)) @ rng.normal(0, 1, (24, D))
    return np.clip(np.rint(128 + 18 * base + rng.normal(0, 6, (229, D))), 0, 255).astype(np.uint8), "synthetic basis"


# ================================================================================================ single GPU
def fit_section(ef, torch):
    """PCA fit seconds (second half of the metric): the shipped Gen-1 shapes through ef_fit_gen1_host (H2D + Gram + Jacobi
    + back-projection + D2H), the oracle's manual_pca (numpy, all host cores) beside it."""
    from oracle import gen1
    out = {}
    X, note = training_matrix()
    cases = [("light", X)]
    dark = os.path.join(GOLDEN, "gen1_dark.npz")
    if os.path.exists(dark):
        cases.append(("dark", np.load(dark)["X_u8"]))
    for name, Xc in cases:
        ef.fit_gen1(Xc, 50)
        torch.cuda.synchronize()
        walls, gpu = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            info = ef.fit_gen1(Xc, 50)[4]
            walls.append(time.perf_counter() - t0)
            gpu.append(info["gpu_ms"])
        cpu = []
        Xf = Xc.astype(np.float64)
        for _ in range(3):
            t0 = time.perf_counter()
            gen1.manual_pca(Xf, 50)
            cpu.append(time.perf_counter() - t0)
        out[name] = {"what": f"manual_pca {Xc.shape[0]}x{Xc.shape[1]} k=50 (tests/golden/gen1_{name}.npz)", "unit": "s",
                     "gpu_wall_s": min(walls), "gpu_device_s": min(gpu) * 1e-3, "jacobi_sweeps": info["sweeps"],
                     "cpu_port_s": min(cpu), "cpu_cores": os.cpu_count()}
    out["note"] = ("gpu_wall_s = ef_fit_gen1_host call incl. H2D/D2H; cpu_port_s = oracle/gen1.py:manual_pca "
                   "(numpy restatement of useless/train.py:56-128, all host cores)")
    return out


def int8_peak_probe(torch, dev):
    """Measured dense int8 tensor throughput of this GPU (torch._int_mm -> cuBLASLt), the denominator for the exact u8
    Gram beside the bf16 figure of MEASURED_PEAKS.json.  None when the op is unavailable."""
    try:
        n = 8192
        a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device=dev)
        ms = min(_time_loop(torch, lambda: torch._int_mm(a, b), 10, 3) for _ in range(3))
        return 2.0 * n ** 3 / ms / 1e9
    except Exception:
        return None


def gram_section(ef, torch, dev, peaks):
    """The exact u8 x u8 tcgen05 SYRK at the config-4 per-GPU shape (12 500 rows x 10 000 pixels -> 10 000 x 10 000)."""
    L = ef._lib.lib()
    N, Dg = 12500, 10000
    x = torch.randint(0, 256, (N, Dg), dtype=torch.uint8, device=dev)
    G = torch.zeros((Dg, Dg), dtype=torch.int64, device=dev)
    wb = int(L.ef_gram_u8_tc_work_bytes(N, Dg, 1))
    work = torch.empty(wb, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run_acc():
        ef._lib.check(L.ef_gram_u8_tc_device(x.data_ptr(), x.stride(0), N, Dg, 0, Dg, 1, G.data_ptr(), work.data_ptr(),
                                             wb, st), "ef_gram_u8_tc_device")

    def run():
        ef._lib.check(L.ef_gram_u8_tc_store_device(x.data_ptr(), x.stride(0), N, Dg, 0, Dg, 1, G.data_ptr(), work.data_ptr(),
                                                   wb, st), "ef_gram_u8_tc_store_device")
    ms_acc = _time_loop(torch, run_acc, 5)
    ms = _time_loop(torch, run, 5)
    # the stored result against the accumulated one (G was zero before the first accumulating call: 7 calls in total)
    Gs = G.clone()
    G.zero_()
    run_acc()
    torch.cuda.synchronize()
    store_equals_accumulate = bool(torch.equal(Gs, G))
    del Gs
    tiles = sum(min((256 * tj + 255) // 128 + 1, (Dg + 127) // 128) for tj in range((Dg + 255) // 256))
    executed = 2.0 * tiles * 128 * 256 * ((N + 127) // 128 * 128)
    algorithmic = 2.0 * Dg * Dg * N / 2.0
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    i8_peak = int8_peak_probe(torch, dev)
    out = {"what": f"ef_gram_u8_tc_store_device side 1: X^T X of u8[{N},{Dg}] -> int64[{Dg},{Dg}] (MN-major operands straight from the row-major X, tcgen05 kind::i8 "
                   "SYRK of the upper-triangle tiles, both triangles stored by the tile epilogues), exact; whole call",
           "ms": ms, "ms_accumulating_call": ms_acc, "store_equals_accumulate": store_equals_accumulate,
           "algorithmic_tops": algorithmic / ms / 1e9, "executed_tops": executed / ms / 1e9,
           "roofline": {"bound": "tensor", "achieved": algorithmic / ms / 1e9, "peak": bf16_peak,
                        "unit": "TOP/s (algorithmic N*D^2 int8 ops, symmetric half) vs measured bf16 TFLOP/s",
                        "frac": algorithmic / ms / 1e9 / bf16_peak,
                        "frac_executed_vs_bf16": executed / ms / 1e9 / bf16_peak,
                        "int8_peak_measured_tops": i8_peak,
                        "frac_executed_vs_int8": (executed / ms / 1e9 / i8_peak) if i8_peak else None,
                        "note": "executed = upper-triangle 128x256 tiles incl. the diagonal overlap; int8 peak = torch._int_mm 8192^3 on this GPU"},
           "flag": int(work[:4].view(torch.int32).item())}
    return out


def preprocess_section(ef, torch, dev, peaks):
    """K1: resize-active preprocess, 4096 square ROIs of 100..300 px inside 1080p frames -> 100x100, gray and BGR."""
    rng = np.random.default_rng(5150)
    F, H, W, nb = 8, 1080, 1920, 4096
    side = rng.integers(100, 301, nb)
    bx = np.stack([rng.integers(0, F, nb), (rng.random(nb) * (W - side)).astype(np.int64),
                   (rng.random(nb) * (H - side)).astype(np.int64), side, side], axis=1).astype(np.int32)
    boxes = torch.from_numpy(bx).to(dev)
    hbm = peaks.get("hbm_gbs", 6650.0)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    out = {}
    for name, shape, ch in (("gray", (F, H, W), 1), ("bgr", (F, H, W, 3), 3)):
        frames = torch.randint(0, 256, shape, dtype=torch.uint8, device=dev)
        outp = ef.engine.preprocess_device(frames, boxes, 100)
        ms = _time_loop(torch, lambda: ef.engine.preprocess_device(frames, boxes, 100, out=outp, bad=bad), 20, 3)
        bytes_alg = float((side.astype(np.int64) ** 2).sum() * ch + nb * 10000)
        out[name] = {"ms": ms, "crops_per_s": nb / ms * 1e3,
                     "roofline": {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                  "frac": bytes_alg / ms / 1e6 / hbm}}
        del frames, outp
    out["what"] = ("ef_preprocess: 4096 square ROIs (100..300 px) of 1080p frames -> 100x100, bit exact cv2.cvtColor + "
                   "cv2.resize INTER_LINEAR; algorithmic bytes = ROI pixels x channels read once + 10 000 B written per crop")
    out["roofline"] = out["gray"]["roofline"]
    out["bad_boxes"] = int(bad.item())
    return out


def template_section(ef, torch, dev):
    """Template-matching detector: one 640x480 frame against 4 persons x 5 crops x 3 scales."""
    rng = np.random.default_rng(640480)
    frame = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    tmpls = [rng.integers(0, 256, (int(rng.integers(80, 121)), int(rng.integers(80, 121))), dtype=np.uint8) for _ in range(20)]
    tmpls[3] = frame[200:300, 250:340].copy()
    matcher = ef.template.TemplateMatcher(tmpls)
    frame_dev = torch.from_numpy(frame).to(dev)
    res = matcher.match(frame_dev)
    torch.cuda.synchronize()
    l0 = ef.launch_count()
    e0, e1 = _events(torch)
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
          "ms_per_frame_device": e0.elapsed_time(e1) / 5, "ms_per_frame_wall": wall * 1e3,
          "launches_per_frame": (ef.launch_count() - l0) / 5,
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
    return tm


SHIPPED = (("gen1_k50_229rows_100x100", 10000, 50, 229, 1, False),
           ("trainv5_k178_178rows_64x64", 4096, 178, 178, 0, True),
           ("trainv4_k50_590rows_64x64", 4096, 50, 590, 0, True),
           ("trainv5_fullk_k272_272rows_64x64", 4096, 272, 272, 0, True),   # train-v5 keeps k = N: the other shipped persons
           ("trainv5_fullk_k308_308rows_64x64", 4096, 308, 308, 0, True),
           ("trainv5_fullk_k590_590rows_64x64", 4096, 590, 590, 0, True))


def shipped_shapes_section(ef, torch, dev):
    """Device-resident recognition time per 4096 crops for the model shapes the reference ships (k = 50 ... 590)."""
    rng = np.random.default_rng(0)
    B = 4096
    out = {"what": "ef_model_recognize_device, 4096 device-resident crops per call, the model shapes the reference ships",
           "unit": "us per 4096 crops"}
    for name, D, k, ng, metric, scaled in SHIPPED:
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        kw = dict(scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D)) if scaled else {}
        rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=metric, **kw)
        ld = (D + 127) // 128 * 128
        xs = [torch.randint(0, 256, (B, ld), dtype=torch.uint8, device=dev) for _ in range(8)]
        res = rec.recognize_device(xs[0], 0.8)
        it = [0]

        def step():
            rec.recognize_device(xs[it[0] % 8], 0.8, out=res)
            it[0] += 1
        l0 = ef.launch_count()
        ms = _time_loop(torch, step, 50, 5)
        out[name] = {"us": ms * 1e3, "crops_per_s": B / ms * 1e3, "launches_per_call": (ef.launch_count() - l0) / 55,
                     "gbs_of_crops": B * D / ms / 1e6}
        rec.close()
        del xs
    return out


def c1_section(ef, torch, dev):
    """BASELINE config 1: train-v5 single person (Joseph_Lai: 178 crops x 4096 pixels, k = N) + recognition of those
    crops.  Ours: ef_fit_gen2_host + Recognizer; CPU: the oracle port of MultiFaceTrainer.train_pca_model
    (train-v5.py:349-385) and of extract_face_features + recognize_face_with_model (scan-template-v4.py:253-287)."""
    from oracle import gen2 as ogen2
    path = os.path.join(GOLDEN, "gen2_joseph.npz")
    if not os.path.exists(path):
        return {"skipped": "tests/golden/gen2_joseph.npz missing"}
    X = np.load(path)["X_u8"]
    N = X.shape[0]
    ef.fit_gen2(X, N)
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        fit = ef.fit_gen2(X, N)
        walls.append(time.perf_counter() - t0)
    labels = np.zeros(N, dtype=np.int32)
    rec = ef.Recognizer(fit["components"], fit["scaler_mean"], fit["features"], scale=fit["scaler_scale"],
                        pca_mean=fit["pca_mean"], labels=labels, metric=ef.METRIC_COSINE_SK, basis_is_components=True)
    r = rec.recognize(X, 0.7)
    t0 = time.perf_counter()
    for _ in range(5):
        r = rec.recognize(X, 0.7, want_features=False)
    batch_s = (time.perf_counter() - t0) / 5
    for i in range(4):                                        # (first launches of the one-crop kernels: lazy module load)
        rec.recognize(X[i:i + 1], 0.7, want_features=False)
    t0 = time.perf_counter()
    for i in range(32):
        rec.recognize(X[i:i + 1], 0.7, want_features=False)
    single_s = (time.perf_counter() - t0) / 32
    cpu_fit = []
    for _ in range(2):
        t0 = time.perf_counter()
        ref = ogen2.train_pca_model(X, N)
        cpu_fit.append(time.perf_counter() - t0)
    m = dict(scaler_mean=ref["scaler_mean"], scaler_scale=ref["scaler_scale"], components=ref["eigenfaces"],
             pca_mean=ref["pca_mean"], face_features=ref["face_features"], face_labels=labels, person_id_map={"Joseph_Lai": 0})
    t0 = time.perf_counter()
    for i in range(32):
        f = ogen2.extract_features(X[i], m["scaler_mean"], m["scaler_scale"], m["components"], m["pca_mean"])[0]
        ogen2.recognize_with_model(f, m["face_features"], m["face_labels"], m["person_id_map"], 0.7)
    cpu_single = (time.perf_counter() - t0) / 32
    t0 = time.perf_counter()
    best, idx, lab = ogen2.recognize_batch(X, m, 0.7)
    cpu_batch = time.perf_counter() - t0
    rec.close()
    return {"what": f"C1: train-v5 single person, {N} crops x {X.shape[1]} pixels, k = {N} (tests/golden/gen2_joseph.npz)",
            "fit": {"gpu_wall_s": min(walls), "gpu_device_s": fit["info"]["gpu_ms"] * 1e-3, "jacobi_sweeps": fit["info"]["sweeps"],
                    "cpu_port_s": min(cpu_fit), "cpu_cores": os.cpu_count()},
            "recognise_the_training_crops": {
                "self_index_exact": bool(np.array_equal(r.index, np.arange(N)) or np.array_equal(r.index, idx)),
                "labels_equal_cpu_port": bool(np.array_equal(r.label, lab)),
                "gpu_batch_crops_per_s": N / batch_s, "gpu_single_crop_ms": single_s * 1e3,
                "cpu_port_batch_crops_per_s": N / cpu_batch, "cpu_port_single_crop_ms": cpu_single * 1e3},
            "note": "host buffers in / results out on both sides; single-crop = one call per crop (the reference's call pattern)"}


def latency_b1_section(ef, torch, dev):
    """The reference's real call pattern: recognize_face_all_models on ONE crop against the 4 shipped person models
    (scan-template-v4.py:289-319; 2.29 ms per model per crop on the survey host).  Microseconds per call."""
    from oracle import gen2 as ogen2
    rng = np.random.default_rng(44)
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    scanner = ef.gen2.MultiModelFaceScanner()
    oracle_models = {}
    for pi, (name, n) in enumerate((("Joseph_Lai", 178), ("ruisheng", 272), ("ruiyi", 308), ("shun", 590))):
        D = 4096
        comps = np.linalg.qr(rng.normal(size=(D, n)))[0].T.copy()
        pca = PCA(n_components=n)
        pca.components_, pca.mean_ = comps, rng.normal(0, 1e-15, D)
        pca.n_components_, pca.n_features_in_, pca.n_samples_ = n, D, n
        pca.explained_variance_ = np.linspace(50, 0.1, n)
        pca.whiten = False
        sc = StandardScaler()
        sc.mean_, sc.scale_ = rng.uniform(60, 200, D), rng.uniform(20, 60, D)
        sc.var_ = sc.scale_ ** 2
        sc.n_features_in_, sc.n_samples_seen_ = D, n
        feats = rng.normal(size=(n, n)) * 30
        md = {"pca": pca, "scaler": sc, "face_features": feats, "face_labels": np.zeros(n, dtype=int),
              "person_id_map": {name: 0}, "n_components": n}
        scanner.models[name] = {"model_data": md, "detection_data": None, "template_images": [], "model_path": ""}
        oracle_models[name] = dict(scaler_mean=sc.mean_, scaler_scale=sc.scale_, components=comps, pca_mean=pca.mean_,
                                   face_features=feats, face_labels=md["face_labels"], person_id_map=md["person_id_map"])
    crop = rng.integers(0, 256, (180, 160, 3), dtype=np.uint8)
    for _ in range(3):
        ours = scanner.recognize_face_all_models(crop, 0.8)
    l0 = ef.launch_count()
    t0 = time.perf_counter()
    n_calls = 50
    for _ in range(n_calls):
        ours = scanner.recognize_face_all_models(crop, 0.8)
    us = (time.perf_counter() - t0) / n_calls * 1e6
    launches = (ef.launch_count() - l0) / n_calls
    from oracle import preprocess as opre
    import cv2
    flat = cv2.resize(cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY), (64, 64)).reshape(1, -1)
    ogen2.recognize_all_models(flat, oracle_models, 0.8)
    t0 = time.perf_counter()
    for _ in range(10):
        flat = cv2.resize(cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY), (64, 64)).reshape(1, -1)
        ref = ogen2.recognize_all_models(flat, oracle_models, 0.8)
    cpu_us = (time.perf_counter() - t0) / 10 * 1e6
    return {"what": "one crop through recognize_face_all_models against 4 person models (k = N = 178 / 272 / 308 / 590), host "
                    "image in, (id, name, confidence) out",
            "gpu_us_per_call": us, "gpu_launches_per_call": launches, "cpu_port_us_per_call": cpu_us,
            "same_answer": bool(ours[1] == ref[1] and abs(ours[2] - ref[2]) < 1e-9),
            "cpu_note": "cv2 gray + resize, then oracle/gen2.py:recognize_all_models (numpy restatement of the sklearn calls, no "
                        "per-call sklearn validation overhead: faster than the reference's own 2.29 ms x 4)"}


def c5_section(ef, torch, dev):
    """BASELINE config 5 on one GPU: synthetic 1080p BGR frames with pasted face crops, host Haar detection (reported
    separately, out of scope), then K1 + K2 for every detected box against 2 person models."""
    import cv2
    path = os.path.join(GOLDEN, "gen2_recog.npz")
    if not os.path.exists(path):
        return {"skipped": "tests/golden/gen2_recog.npz missing"}
    g = np.load(path)
    crops = [g[f"crop_{i:02d}"] for i in range(int(g["n_crops"]))]
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    scanner = ef.gen2.MultiModelFaceScanner()
    for name in [str(p) for p in g["persons"]]:
        comps = g[f"{name}_components"]
        k, D = comps.shape
        pca = PCA(n_components=k)
        pca.components_, pca.mean_ = comps, g[f"{name}_pca_mean"]
        pca.n_components_, pca.n_features_in_ = k, D
        pca.explained_variance_ = g[f"{name}_explained_variance"]
        sc = StandardScaler()
        sc.mean_, sc.scale_, sc.var_ = g[f"{name}_scaler_mean"], g[f"{name}_scaler_scale"], g[f"{name}_scaler_var"]
        sc.n_features_in_ = D
        md = {"pca": pca, "scaler": sc, "face_features": g[f"{name}_face_features"],
              "face_labels": g[f"{name}_face_labels"], "person_id_map": {name: 0}, "n_components": k}
        scanner.models[name] = {"model_data": md, "detection_data": None, "template_images": [], "model_path": ""}
    rng = np.random.default_rng(5150)
    n_frames = 12
    frames, boxes_per_frame = [], []
    cascade = ef.pipeline.haar_detector()
    t_haar = 0.0
    for _ in range(n_frames):
        frame = cv2.GaussianBlur(rng.integers(90, 166, (1080, 1920, 3), dtype=np.uint8), (0, 0), 3)
        for _ in range(int(rng.integers(1, 5))):
            c = crops[int(rng.integers(0, len(crops)))]
            if c.ndim == 2:
                c = cv2.cvtColor(c, cv2.COLOR_GRAY2BGR)
            s = int(rng.integers(140, 320))
            c = cv2.resize(c, (s, s))
            x, y = int(rng.integers(0, 1920 - s)), int(rng.integers(0, 1080 - s))
            frame[y:y + s, x:x + s] = c
        t0 = time.perf_counter()
        bx = ef.pipeline.detect_boxes(cascade, frame)
        t_haar += time.perf_counter() - t0
        frames.append(frame)
        boxes_per_frame.append(bx)
    n_boxes = sum(len(b) for b in boxes_per_frame)
    for f, b in zip(frames[:2], boxes_per_frame[:2]):
        ef.pipeline.recognize_frame(scanner, f, b, 0.8)
    t0 = time.perf_counter()
    for f, b in zip(frames, boxes_per_frame):
        ef.pipeline.recognize_frame(scanner, f, b, 0.8)
    t_gpu = time.perf_counter() - t0
    # all frames' boxes in one batch (frames resident on the device): the data-parallel unit a rank would process
    allb = np.concatenate([np.concatenate([np.full((len(b), 1), i, np.int32), b], axis=1)
                           for i, b in enumerate(boxes_per_frame) if len(b)] or [np.zeros((0, 5), np.int32)])
    out = {"what": f"C5 (one GPU): {n_frames} synthetic 1080p BGR frames with pasted golden face crops; host Haar "
                   "detectMultiScale(1.1, 5, (30, 30)) -> boxes -> K1 (BGR2GRAY + resize 64x64) + K2 against 2 person models",
           "frames": n_frames, "boxes_detected": n_boxes,
           "host_haar_ms_per_frame": t_haar / n_frames * 1e3,
           "gpu_recognise_ms_per_frame_host_frames": t_gpu / n_frames * 1e3,
           "frames_per_s_end_to_end": n_frames / (t_haar + t_gpu),
           "note": "per-frame call uploads the 6.2 MB frame (host BGR) every time; Haar on the host dominates (out of scope per "
                   "BASELINE.json north_star)"}
    if len(allb):
        stack = torch.from_numpy(np.stack(frames)).to(dev)
        bdev = torch.from_numpy(np.ascontiguousarray(allb)).to(dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        recs = [ef.gen2.recognizer_for(info["model_data"]) for info in scanner.models.values()]

        def batch():
            x = ef.engine.preprocess_device(stack, bdev, 64, bad=bad)
            for rec in recs:
                rec.recognize_device(x, 0.8, want_residual=False)
        ms = _time_loop(torch, batch, 10, 2)
        out["device_resident_batch"] = {"ms_per_batch_of_all_boxes": ms, "boxes_per_s": len(allb) / ms * 1e3,
                                        "what": "frames already on the device, all boxes of all frames: one K1 launch + K2 per model"}
    return out


# ================================================================================================ multi GPU
def _c3_data(torch, dev, n3, k3, B3):
    gen = torch.Generator(device=dev)
    gen.manual_seed(1_000_003)
    lam3 = torch.arange(1, k3 + 1, device=dev, dtype=torch.float64) ** -2.0
    G3 = torch.randn((n3, k3), generator=gen, device=dev, dtype=torch.float64) * lam3.sqrt()
    truth = torch.randint(0, n3, (B3,), generator=gen, device=dev)
    P3 = G3[truth] + 0.05 * torch.randn((B3, k3), generator=gen, device=dev, dtype=torch.float64) * lam3.sqrt()
    return G3, truth, P3


def large_gallery_sharded(ef, torch, dist, dev, rank, world, peaks, cpu_leg=True, n3=1_000_000, k3=128, B3=4096):
    """BASELINE config 3 at full size: 1 M gallery identities x k = 128 split into contiguous row shards over the ranks
    (dist.shard_bounds), 4096 queries replicated; per rank ef_match_tc_device on its shard (tcgen05 f16 hi/lo filter +
    exact float64 re-score), ONE NCCL all-gather of (score, global index) and ef_match_reduce_device.  Strong scaling.
    Parity inside the run: rank 0 also matches against the WHOLE gallery on its own GPU and compares bit for bit."""
    G3, truth, P3 = _c3_data(torch, dev, n3, k3, B3)         # same generator seed on every rank: identical data
    lo, hi = ef.dist.shard_bounds(n3, world, rank)
    sg = ef.dist.ShardedGallery(G3[lo:hi], lo, ef.METRIC_COSINE_SK)
    for _ in range(2):
        sg.match(P3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = _events(torch)
    phases = {}
    e0.record()
    for _ in range(reps):
        s3, i3 = sg.match(P3, timings=False)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1) / reps)
    for _ in range(3):                                      # phase split on separate repetitions (the stopwatch synchronises)
        sg.match(P3, timings=True)
        for kname, v in sg.last_timings.items():
            phases[kname] = phases.get(kname, 0.0) + v * 1e3 / 3
    phases = {kname: _max_over_ranks(torch, dist, world, dev, v) for kname, v in sorted(phases.items())}
    # the same shards under the Euclidean metric (the reference's first matcher variant): one extra GEMM component
    sg_l2 = ef.dist.ShardedGallery(G3[lo:hi], lo, ef.METRIC_L2)
    for _ in range(2):
        sg_l2.match(P3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        s3e, i3e = sg_l2.match(P3, timings=False)
    e1.record()
    torch.cuda.synchronize()
    ms_l2 = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1) / reps)
    l2_flags = sg_l2.last_flags
    del sg_l2
    out = None
    identical, score_diff, acc = None, None, None
    if rank == 0:
        sg64 = ef.dist.ShardedGallery(G3, 0, ef.METRIC_L2, use_tensor_cores=False)
        s64, i64 = sg64.match_local(P3[:64])
        l2 = {"ms_per_batch": ms_l2, "queries_per_s": B3 / ms_l2 * 1e3,
              "bit_identical_to_float64_scan_on_64_queries": bool(torch.equal(i64, i3e[:64]) and torch.equal(s64, s3e[:64])),
              "top1_accuracy_vs_planted": float((i3e == truth).double().mean()),
              "candidates_rescored_on_rank0_shard": l2_flags["candidates"] if l2_flags else None}
        del sg64
        whole = ef.dist.ShardedGallery(G3, 0, ef.METRIC_COSINE_SK) if world > 1 else sg
        ws, wi = whole.match_local(P3)
        identical = bool(torch.equal(wi, i3) and torch.equal(ws, s3))
        score_diff = float((ws - s3).abs().max())
        acc = float((i3 == truth).double().mean())
        # float64 CUDA-core scan on 64 queries: the filter path must agree bit for bit with it as well
        sg64 = ef.dist.ShardedGallery(G3, 0, ef.METRIC_COSINE_SK, use_tensor_cores=False)
        s64, i64 = sg64.match_local(P3[:64])
        scan_identical = bool(torch.equal(i64, i3[:64]) and torch.equal(s64, s3[:64]))
        del sg64
        if whole is not sg:
            del whole
    if world > 1:
        dist.barrier()
    if rank == 0:
        bf16_peak = peaks.get("bf16_tflops", 1590.0)
        alg_flops = 2.0 * B3 * n3 * k3                       # 2 * N_g * k per query
        out = {"what": f"C3: {B3} queries x {n3} gallery rows x k = {k3}, cosine top-1, gallery sharded by rows over {world} "
                       "GPU(s): ef_match_tc_device per shard + NCCL all-gather of (score, index) + ef_match_reduce_device",
               "n_gpus": world, "scaling": "strong", "ms_per_batch": ms, "queries_per_s": B3 / ms * 1e3,
               "phases_ms_max_over_ranks": phases,
               "collective_share": (phases.get("allgather", 0.0) + phases.get("reduce", 0.0)) / max(phases.get("total", ms), 1e-9),
               "bit_identical_to_single_shard": identical, "queries_compared": B3, "score_max_abs_diff": score_diff,
               "bit_identical_to_float64_scan_on_64_queries": scan_identical,
               "top1_accuracy_vs_planted": acc, "candidates_rescored_on_rank0_shard": sg.last_flags["candidates"] if sg.last_flags else None,
               "euclidean_metric": l2,
               "roofline": {"bound": "tensor", "achieved": alg_flops / ms / 1e9, "peak": bf16_peak * world,
                            "unit": "TFLOP/s algorithmic (2*B*Ng*k) vs measured bf16 peak x GPUs",
                            "frac": alg_flops / ms / 1e9 / (bf16_peak * world)}}
        if cpu_leg:
            Gh = G3.cpu().numpy()
            Ph = P3[:256].cpu().numpy()
            legs = {}
            for dt in (np.float64, np.float32):
                Gd = Gh.astype(dt)
                Gd /= np.maximum(np.linalg.norm(Gd, axis=1, keepdims=True), 1e-300).astype(dt)
                Pd = Ph.astype(dt)
                t0 = time.perf_counter()
                best = np.full(len(Pd), -np.inf)
                arg = np.zeros(len(Pd), dtype=np.int64)
                for j in range(0, n3, 65536):
                    sblk = Pd @ Gd[j:j + 65536].T
                    a = sblk.argmax(1)
                    v = sblk[np.arange(len(Pd)), a]
                    upd = v > best
                    best[upd], arg[upd] = v[upd], a[upd] + j
                dt_s = time.perf_counter() - t0
                legs[np.dtype(dt).name] = {"seconds_256_queries": dt_s, "queries_per_s": 256 / dt_s,
                                           "ms_per_4096_batch_extrapolated": dt_s * 16 * 1e3,
                                           "argmax_equal_gpu": bool(np.array_equal(arg, i3[:256].cpu().numpy()))}
            out["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "legs": legs,
                                   "sample": "256 of the 4096 queries, numpy blocked Q @ G.T + running argmax (BASELINE.md section 3, "
                                             "C3), gallery normalised once outside the timed region; extrapolated x16"}
    del G3, P3, sg
    torch.cuda.empty_cache()
    return out


def _c4_rows(torch, dev, lo, hi, D4, R4, F4, sig4, chunk=5000):
    """Rows [lo, hi) of the planted-factor matrix of SURVEY 8(d) C4; chunk c always comes from seed 4243 + c, so any
    partition of the rows over ranks yields the same global matrix."""
    X = torch.empty((hi - lo, D4), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    for c in range(lo // chunk, (hi + chunk - 1) // chunk):
        gen.manual_seed(4243 + c)
        L4 = torch.randn((chunk, R4), generator=gen, device=dev) * sig4
        blk = (128 + L4 @ F4.T + 4.0 * torch.randn((chunk, D4), generator=gen, device=dev)).round_().clamp_(0, 255).to(torch.uint8)
        a, b = max(lo, c * chunk), min(hi, (c + 1) * chunk)
        X[a - lo:b - lo] = blk[a - c * chunk:b - c * chunk]
    return X


def fit_sharded(ef, torch, dist, dev, rank, world, peaks, cpu_leg=True, full_cpu=False, N4=100_000, D4=10_000, K4=256):
    """BASELINE config 4: manual_pca's covariance branch (useless/train.py:97-122) on N = 100 000 synthetic faces x
    D = 10 000 pixels, k = 256, rows sharded over the ranks: per rank exact integer column sums + tcgen05 integer Gram of its
    rows, ONE NCCL all-reduce(SUM) of int64 [D*D + D], exact centring, top-k eigenpairs (filtered subspace iteration, the
    covariance products sharded by rows with an all-gather each), local projection.  Strong scaling (N fixed)."""
    R4 = 300
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    F4 = torch.linalg.qr(torch.randn((D4, R4), generator=gen, device=dev, dtype=torch.float32))[0]
    sig4 = 40.0 * torch.arange(1, R4 + 1, device=dev, dtype=torch.float32) ** -0.7
    lo, hi = ef.dist.shard_bounds(N4, world, rank)
    X = _c4_rows(torch, dev, lo, hi, D4, R4, F4, sig4)
    solos = [dist.new_group([r]) for r in range(world)] if world > 1 else None
    ef.dist.fit_gen1_sharded(X, N4, K4)                      # warm-up (allocations, NCCL channels)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E4, mean4, proj4, ev4 = ef.dist.fit_gen1_sharded(X, N4, K4, timings=True)
    torch.cuda.synchronize()
    wall = _max_over_ranks(torch, dist, world, dev, time.perf_counter() - t0)
    tm = dict(ef.dist.fit_gen1_sharded.last_timings)
    tm = {kname: _max_over_ranks(torch, dist, world, dev, v) for kname, v in sorted(tm.items())}
    info4 = dict(getattr(ef.dist.fit_gen1_sharded, "last_solver_info", {}))
    same_everywhere = True
    if world > 1:
        gathered = [torch.empty_like(ev4) for _ in range(world)]
        dist.all_gather(gathered, ev4)
        same_everywhere = all(bool(torch.equal(gg, ev4)) for gg in gathered)
    out = None
    single = {}
    if rank == 0 and world > 1:
        Xall = _c4_rows(torch, dev, 0, N4, D4, R4, F4, sig4)
        E1, mean1, _, ev1 = ef.dist.fit_gen1_sharded(Xall, N4, K4, group=solos[0])
        single = {"eigenvalues_bit_identical_to_1_rank": bool(torch.equal(ev1, ev4)),
                  "eigenvalues_max_rel_diff": float(((ev1 - ev4).abs() / ev1.abs().clamp_min(1e-300)).max()),
                  "mean_bit_identical": bool(torch.equal(mean1, mean4)),
                  "eigenvector_min_abs_cos": float((E1 * E4).sum(0).abs().min())}
        del Xall, E1
    if world > 1:
        dist.barrier()
    if rank == 0:
        top = 32
        # principal angles between the top eigenvectors and the planted factor subspace (float64 SVD of a 32 x 300 matrix)
        cosines = torch.linalg.svdvals(E4[:, :top].T @ F4.double())
        gram_flops = float(N4) * D4 * D4                    # symmetric half, int8 MACs x 2 / 2
        bf16_peak = peaks.get("bf16_tflops", 1590.0)
        out = {"what": f"C4: manual_pca covariance branch on u8[{N4},{D4}] (planted rank-{R4} factors), k = {K4}, rows sharded "
                       f"over {world} GPU(s) ({hi - lo} rows on rank 0)",
               "n_gpus": world, "scaling": "strong", "seconds": wall, "phases_s_max_over_ranks": tm,
               "collective": {"allreduce_int64_bytes": (D4 * D4 + D4) * 8, "allreduce_s": tm.get("allreduce"),
                              "share_of_total": (tm.get("allreduce", 0.0) / tm["total"]) if tm.get("total") else None,
                              "solver_allgathers": info4.get("allgathers", 0)},
               "solver": info4, "eigenvalues_identical_on_every_rank": same_everywhere,
               "orthonormality_error": float((E4.T @ E4 - torch.eye(K4, device=dev, dtype=torch.float64)).abs().max()),
               "min_cos_principal_angle_top32_vs_planted": float(cosines.min()),
               "gram_roofline": {"bound": "tensor", "achieved": gram_flops / max(tm.get("gram", 1e-9), 1e-9) / 1e12,
                                 "peak": bf16_peak * world, "unit": "TOP/s algorithmic (N*D^2) vs measured bf16 peak x GPUs",
                                 "frac": gram_flops / max(tm.get("gram", 1e-9), 1e-9) / 1e12 / (bf16_peak * world)}}
        out.update(single)
        if cpu_leg:
            Ns, ne = (10_000, D4) if full_cpu else (2_000, 2_500)
            Xs = X[:Ns].cpu().numpy().astype(np.float64)
            t0 = time.perf_counter()
            Xc = Xs - Xs.mean(0)
            cov = np.cov(Xc.T)                               # useless/train.py:99
            t_cov = time.perf_counter() - t0
            t0 = time.perf_counter()
            np.linalg.eigh(cov[:ne, :ne])                    # :103
            t_eigh = time.perf_counter() - t0
            out["cpu_baseline"] = {
                "kind": "port", "cores": os.cpu_count(),
                "np_cov_seconds": t_cov, "np_cov_rows": Ns, "np_cov_seconds_extrapolated_full_N": t_cov * N4 / Ns,
                "eigh_seconds": t_eigh, "eigh_n": ne, "eigh_seconds_extrapolated_n10000": t_eigh * (D4 / ne) ** 3,
                "seconds_extrapolated": t_cov * N4 / Ns + t_eigh * (D4 / ne) ** 3,
                "sample": f"reference manual_pca else-branch (np.cov + np.linalg.eigh, useless/train.py:97-103): np.cov on {Ns} of the "
                          f"{N4} rows (extrapolated linearly in N), eigh on the leading {ne} x {ne} block (extrapolated with n^3); "
                          "--full-cpu-legs times N = 10 000 rows and the full 10 000 x 10 000 eigh"}
    del X, E4, proj4
    torch.cuda.empty_cache()
    return out


def h2d_concurrency(torch, dist, dev, world, nbytes=40_960_000, reps=20):
    """Why the end-to-end number scales sub-linearly: every rank uploads the bench's 41 MB batch from page-locked host memory
    at the same time, nothing else running.  Aggregate GB/s over all ranks (max-over-ranks time)."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1) / reps)
    return {"what": "concurrent pinned host -> device copies of one 41 MB batch per rank, no kernels",
            "ms_per_copy_max_over_ranks": ms, "aggregate_gbs": world * nbytes / ms / 1e6,
            "per_gpu_gbs": nbytes / ms / 1e6}
