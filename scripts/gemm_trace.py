"""Per-phase timestamps of the skinny GEMM (debug entry dflash_gemm_trace): where a launch's time goes as the
activation width grows.  python scripts/gemm_trace.py   (GPU box)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dflash_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.dflash_gemm_trace.restype = ctypes.c_int
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
torch.manual_seed(0)
names = ["entry", "prologue", "pdl_wait", "first_stage", "last_mma_issued", "last_acc_done", "stores_issued"]
for (N, K, label) in [(4096, 4096, "o"), (6144, 4096, "qkv"), (24576, 4096, "gate/up"), (4096, 12288, "down")]:
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    for mb in (16, 128, 256):
        X = torch.randn(mb, K, device="cuda").to(torch.bfloat16)
        slots = lib.dflash_gemm_max_slots(N, K, 148)
        ws = torch.zeros(slots, mb, N, dtype=torch.float32, device="cuda")
        tr = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        best = None
        for rep in range(4):
            flush.zero_()  # weights out of L2
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g = lib.dflash_gemm_trace(P(W), N, K, P(X), mb, mb, mb, P(ws), mb, P(tr), 148, st)
            e1.record()
            torch.cuda.synchronize()
            assert g > 0, g
            t = tr.view(148, 8)[:g].cpu().double()
            t0 = t[:, 0].min()
            rel = (t[:, :7] - t0) / 1e3
            best = (e0.elapsed_time(e1) * 1e3, rel)
        us, rel = best
        med = rel.median(dim=0).values.tolist()
        mx = rel.max(dim=0).values.tolist()
        print(f"{label:8s} N={N} K={K} mb={mb:3d}: {us:6.1f} us | median per-CTA us since first entry: " +
              " ".join(f"{n}={v:.1f}" for n, v in zip(names, med)) + f" | max stores_issued={mx[6]:.1f}", flush=True)
