"""Micro-benchmark of the tensor-core Gram (ef_gram_u8_tc_device): device time by CUDA events, int8 TOP/s.  Not a bench line."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

L = ef._lib.lib()
shapes = [(229, 10000, 0), (590, 4096, 0), (4096, 10000, 0), (12500, 4096, 1), (12500, 10000, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for N, D, side in shapes:
    x = torch.randint(0, 256, (N, (D + 15) // 16 * 16), dtype=torch.uint8, device="cuda")[:, :D]
    n, K = (N, D) if side == 0 else (D, N)
    G = torch.zeros((n, n), dtype=torch.int64, device="cuda")
    wb = int(L.ef_gram_u8_tc_work_bytes(N, D, side))
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        fn = L.ef_gram_u8_tc_store_device if os.environ.get("EF_GRAM_PROBE_STORE") else L.ef_gram_u8_tc_device
        ef._lib.check(fn(x.data_ptr(), x.stride(0), N, D, 0, D, side, G.data_ptr(), work.data_ptr(), wb, st), "gram")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    reps = 5 if n > 2000 else 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    full = 2.0 * n * n * K
    print(f"N={N} D={D} side={side}: n={n} K={K}  {ms*1e3:9.1f} us   {full/ms/1e9:8.1f} TOP/s (full n^2 K count; upper "
          f"triangle only is computed)  flag={int(work[:4].view(torch.int32).item())}", flush=True)
