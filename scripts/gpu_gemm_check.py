"""GPU bring-up check for the tcgen05 skinny GEMM: correctness vs torch and achieved GB/s.

Run on the GPU box:  python scripts/gpu_gemm_check.py
"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dflash_b200 import _lib  # noqa: E402


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def run_gemm(lib, W, X, mb, m_valid, grid, pdl=0, w_row0=0, N=None):
    Ntot, K = W.shape
    N = N or Ntot
    slots = lib.dflash_gemm_max_slots(N, K, grid)
    assert slots > 0, slots
    ws = torch.zeros(slots, mb, N, dtype=torch.float32, device="cuda")
    out = torch.zeros(m_valid, N, dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.dflash_gemm_skinny(ptr(W), Ntot, w_row0, N, K, ptr(X), X.shape[0], 0, mb, m_valid, ptr(ws), mb,
                                N, ptr(out), N, grid, pdl, st)
    _lib.check(rc, "gemm_skinny")
    return out, ws


def main():
    lib = _lib.load()
    sms = _lib.check(lib.dflash_device_check(), "device_check")
    print("SMs", sms, flush=True)
    torch.manual_seed(0)
    ok = True
    cases = [(256, 128, 16), (4096, 4096, 16), (6144, 4096, 32), (4096, 12288, 16), (24576, 4096, 16),
             (4096, 20480, 16), (1000, 512, 16), (4096, 4096, 64), (4096, 4096, 128), (4096, 4096, 256)]
    for (N, K, mb) in cases:
        W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        X = torch.randn(mb, K, device="cuda").to(torch.bfloat16)
        for grid in (sms, 2 * sms, 7):
            out, _ = run_gemm(lib, W, X, mb, mb, grid)
            torch.cuda.synchronize()
            ref = X.float() @ W.float().t()
            err = (out - ref).abs().max().item()
            scale = ref.abs().max().item()
            good = err <= 2e-3 * max(scale, 1.0)
            ok &= good
            print(f"N={N} K={K} mb={mb} grid={grid}: max|err|={err:.3e} (ref max {scale:.3f}) {'OK' if good else 'FAIL'}",
                  flush=True)
    # weight sub-range
    W = (torch.randn(6144, 4096, device="cuda") * 0.05).to(torch.bfloat16)
    X = torch.randn(16, 4096, device="cuda").to(torch.bfloat16)
    out, _ = run_gemm(lib, W, X, 16, 16, sms, w_row0=4096, N=2048)
    ref = X.float() @ W[4096:].float().t()
    err = (out - ref).abs().max().item()
    print("sub-range err", err)
    ok &= err < 2e-2

    # argmax mode
    V, K = 151936, 4096
    W = (torch.randn(V, K, device="cuda") * 0.02).to(torch.bfloat16)
    X = torch.randn(16, K, device="cuda").to(torch.bfloat16)
    cand_v = torch.empty(sms, 16, dtype=torch.float32, device="cuda")
    cand_i = torch.empty(sms, 16, dtype=torch.int32, device="cuda")
    logits = torch.empty(16, V, dtype=torch.bfloat16, device="cuda")
    toks = torch.empty(16, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dflash_gemm_argmax(ptr(W), V, V, K, ptr(X), 16, 0, 16, 16, ptr(cand_v), ptr(cand_i),
                                      ptr(logits), V, ptr(toks), sms, 0, st), "gemm_argmax")
    torch.cuda.synchronize()
    ref_logits = (X.float() @ W.float().t())
    ref_b = ref_logits.to(torch.bfloat16)
    lerr = (logits.float() - ref_logits).abs().max().item()
    print("lm_head logits max err", lerr, "argmax equal to argmax(own logits):",
          bool((toks == logits.float().argmax(-1)).all()), "equal to torch ref:",
          int((toks == ref_b.float().argmax(-1)).sum()), "/16")
    ok &= bool((toks == logits.float().argmax(-1)).all())

    # timing: stream the weights
    def bench(N, K, mb, grid, pdl, iters=20, argmax=False):
        W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        X = torch.randn(mb, K, device="cuda").to(torch.bfloat16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        slots = lib.dflash_gemm_max_slots(N, K, grid)
        ws = torch.zeros(slots, mb, N, dtype=torch.float32, device="cuda")
        out = torch.zeros(mb, N, dtype=torch.float32, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts = []
        for i in range(iters + 3):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if argmax:
                lib.dflash_gemm_argmax(ptr(W), N, N, K, ptr(X), mb, 0, mb, mb, ptr(cand_v), ptr(cand_i), None,
                                       0, ptr(toks), grid, pdl, st)
            else:
                lib.dflash_gemm_skinny(ptr(W), N, 0, N, K, ptr(X), mb, 0, mb, mb, ptr(ws), mb, N, ptr(out), N,
                                       grid, pdl, st)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        gbs = N * K * 2 / med / 1e6
        print(f"bench N={N} K={K} mb={mb} grid={grid} pdl={pdl} argmax={argmax}: median {med*1000:.1f} us "
              f"-> {gbs:.0f} GB/s (incl. tiny reduce kernel)", flush=True)

    for grid in (sms, 2 * sms):
        bench(24576, 4096, 16, grid, 0)
        bench(4096, 12288, 16, grid, 0)
        bench(4096, 4096, 16, grid, 0)
        bench(6144, 4096, 32, grid, 0)
        bench(4096, 20480, 16, grid, 0)
    bench(151936, 4096, 16, sms, 0, argmax=True)
    bench(24576, 4096, 64, sms, 0)
    bench(24576, 4096, 256, sms, 0)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
