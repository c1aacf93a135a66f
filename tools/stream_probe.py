"""Debugging aid for the persistent serving kernel (recognize_stream_kernel) on the bench shape (C2): steady-state time per
4096-crop batch for several queue depths and shared-memory plans, the pipelined kernel of round 1 beside it, the in-kernel
phase stamps (EF_TC_PROBE) and the host cost of one submit call.  Not a bench line."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B, D, k, ng = 4096, 10000, 10, 1024
rng = np.random.default_rng(0)
E = np.linalg.qr(rng.normal(size=(D, k)))[0]
rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=ef.METRIC_COSINE_G1,
                    labels=np.arange(ng) % 4)
ld = (D + 127) // 128 * 128
xs = [torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda") for _ in range(8)]
outs = [rec.recognize_device(x, 0.8) for x in xs]
want = [{f: v.clone() for f, v in o.items()} for o in outs]
torch.cuda.synchronize()


def run(steps):
    for i in range(steps):
        rec.submit_device(xs[i % 8], 0.8, out=outs[i % 8])
    rec.flush_device()


def timed(steps=200, reps=5):
    run(40)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    return best * 1e3


def check():
    for o, w in zip(outs, want):
        for f in ("features", "score", "index", "label", "resid2"):
            if not torch.equal(o[f], w[f]):
                return f"MISMATCH in {f}: {(o[f] != w[f]).sum().item()} of {o[f].numel()}"
    return "bit identical"


algo = (D + 4 * k + 8) * B
configs = [({}, "default plan")]
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    configs += [({"EF_STREAM_RESIDENT": "0"}, "gallery ring"), ({"EF_STREAM_STAGES": "3"}, "3 stages"),
                ({"EF_STREAM_STAGES": "3", "EF_STREAM_RECV_BUFS": "2"}, "3 stages, 2 receive buffers")]
for env, name in configs:
    for kk in ("EF_STREAM_RESIDENT", "EF_STREAM_STAGES", "EF_STREAM_RECV_BUFS"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    for depth in (16, -16, -8, -4, -2, -1):          # 16 = adaptive (default); negative = fixed depth
        rec.set_serving(0, depth)
        for o in outs:
            for name, v in o.items():
                if v is not None and not name.startswith("_"):
                    v.zero_()
        us = timed()
        print(f"stream kernel [{name}] depth {depth if depth > 0 else str(-depth) + ' fixed'}: {us:7.2f} us per batch = {algo / us / 1e3:7.1f} GB/s = "
              f"{algo / us / 1e3 / 6550.1:.3f} of HBM peak; {check()}; timeouts {rec.pipeline_timeouts()}", flush=True)
for kk in ("EF_STREAM_RESIDENT", "EF_STREAM_STAGES", "EF_STREAM_RECV_BUFS"):
    os.environ.pop(kk, None)
if len(sys.argv) > 1 and sys.argv[1] == "l2":
    # the same 41 MB batch every time: the crops come from L2 (126 MB) instead of HBM
    rec.set_serving(0, 8)
    os.environ["EF_STREAM_PREFETCH"] = "0"
    keep = xs
    print(f"stream kernel, 8 distinct batches (HBM), depth 8: {timed():7.2f} us per batch", flush=True)
    xs = [keep[0]] * 8
    print(f"stream kernel, ONE batch re-used (L2 resident), depth 8: {timed():7.2f} us per batch", flush=True)
    half = [keep[0][:2048], keep[1][:2048]] * 4
    xs = half
    outs_keep = outs
    outs = [{f: (v[:2048] if v is not None else None) for f, v in o.items() if not f.startswith("_")} for o in outs]
    print(f"stream kernel, 2048-crop batches (16 clusters = 64 SMs), depth 8: {timed():7.2f} us per batch", flush=True)
    xs, outs = keep, outs_keep
    os.environ.pop("EF_STREAM_PREFETCH")
if len(sys.argv) > 1 and sys.argv[1] == "stages":
    rec.set_serving(0, 8)
    for st in (2, 3, 4):
        os.environ["EF_STREAM_STAGES"] = str(st)
        us = timed()
        print(f"stream kernel, {st} stages, depth 8: {us:7.2f} us per batch = {algo / us / 1e3 / 6550.1:.3f} of HBM peak; {check()}", flush=True)
    os.environ.pop("EF_STREAM_STAGES")
if len(sys.argv) > 1 and sys.argv[1] == "xbox":
    rec.set_serving(0, 8)
    for xb in (128, 64, 32, 16, 8):
        os.environ["EF_STREAM_XBOX"] = str(xb)
        us = timed()
        print(f"stream kernel, crop tile as {128 // xb} TMA boxes of {xb} rows, depth 8: {us:7.2f} us per batch = {algo / us / 1e3 / 6550.1:.3f} of HBM peak; {check()}", flush=True)
    os.environ.pop("EF_STREAM_XBOX")
if len(sys.argv) > 1 and sys.argv[1] == "prefetch":
    rec.set_serving(0, 8)
    for pf in (0, 4, 8, 12, 16, 24, 40):
        os.environ["EF_STREAM_PREFETCH"] = str(pf)
        us = timed()
        print(f"stream kernel, L2 prefetch {pf:2d} K blocks ahead, depth 8: {us:7.2f} us per batch = {algo / us / 1e3 / 6550.1:.3f} of HBM peak; {check()}", flush=True)
    os.environ.pop("EF_STREAM_PREFETCH")
if len(sys.argv) > 1 and sys.argv[1] == "debug":
    # interference matrix: one role's work switched off at a time (EF_STREAM_DEBUG bits; the results are then wrong)
    rec.set_serving(0, -16)
    names = {0: "all roles", 1: "no sum-of-squares reads", 2: "drain: tcgen05.ld only (no combine-to-int64, no DSMEM push)",
             6: "drain: nothing", 8: "combine: no integer sums", 16: "scan: no tcgen05.ld of the scores", 32: "re-score: nothing",
             64: "filter: no MMAs", 126: "everything but the stream off", 127: "everything off"}
    for bits, nm in names.items():
        os.environ["EF_STREAM_DEBUG"] = str(bits)
        print(f"stream kernel, depth 16 fixed, debug {bits:3d} ({nm}): {timed():7.2f} us per batch; timeouts {rec.pipeline_timeouts()}", flush=True)
    os.environ.pop("EF_STREAM_DEBUG")
    os.environ["EF_TC_PROBE"] = "1"
    rec.set_serving(0, -8)
    run(8)
    torch.cuda.synchronize()
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "limits":
    # where does the stream time go?  (a) without the sum-of-squares reads of the staged tiles (results then differ),
    # (b) with half the digit planes (S = 4: half the basis traffic and half the MMA operand reads)
    rec.set_serving(0, 8)
    os.environ["EF_STREAM_NO_SSQ"] = "1"
    print(f"stream kernel, no sum-of-squares reads, depth 8: {timed():7.2f} us per batch", flush=True)
    os.environ.pop("EF_STREAM_NO_SSQ")
    rec4 = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=ef.METRIC_COSINE_G1,
                         labels=np.arange(ng) % 4, n_slices=4)
    o4 = [rec4.recognize_device(x, 0.8) for x in xs]
    keep, rec = rec, rec4
    outs_keep, outs = outs, o4
    rec.set_serving(0, 8)
    print(f"stream kernel, S = 4 digit planes, depth 8: {timed():7.2f} us per batch", flush=True)
    os.environ["EF_STREAM_NO_SSQ"] = "1"
    print(f"stream kernel, S = 4, no sum-of-squares reads, depth 8: {timed():7.2f} us per batch", flush=True)
    os.environ.pop("EF_STREAM_NO_SSQ")
    rec, outs = keep, outs_keep
rec.set_serving(1, 0)
us = timed()
print(f"pipelined kernel (round 1): {us:7.2f} us per batch = {algo / us / 1e3 / 6550.1:.3f} of HBM peak; {check()}", flush=True)
rec.set_serving(0, 8)
# host cost of a submit call (python + ctypes + descriptor encode), nothing launched until the queue fills
torch.cuda.synchronize()
t0 = time.perf_counter()
run(800)
host_us = (time.perf_counter() - t0) / 800 * 1e6
torch.cuda.synchronize()
print(f"host time per submit_device call (incl. 1 launch per 8): {host_us:.2f} us", flush=True)
os.environ["EF_TC_PROBE"] = "1"
run(8)
torch.cuda.synchronize()
run(8)
torch.cuda.synchronize()
os.environ.pop("EF_TC_PROBE")
