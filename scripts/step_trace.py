"""In-graph timeline of one draft+verify step (debug build):
    nvcc ... -DDFLASH_STEP_TRACE -o build/lib_trace.so dflash_b200/csrc/api.cu -lcudart
    DFLASH_LIB=$PWD/build/lib_trace.so python scripts/step_trace.py
Thread 0 of block 0 of every kernel stamps globaltimer at entry (phase 0), after griddepcontrol.wait (1) and at its
end (2); GEMMs also when their first stage has landed (3). Kernels are recognised by launch shape."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
dims = bench.Q8
bs, H, V, L = dims["block_size"], dims["hidden"], dims["vocab"], dims["draft_layers"]
draft, eng, embed, lm_head = bench.build_engine(dims, dev, seed=0)
lib = eng.lib
lib.dflash_step_trace_set.restype = ctypes.c_int
lib.dflash_step_trace_set.argtypes = [ctypes.c_void_p, ctypes.c_uint]
g = torch.Generator(device=dev).manual_seed(100)
tl = torch.randn(bs, V, device=dev, generator=g).to(torch.bfloat16)
hsel = [(torch.randn(bs, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
ph = [(torch.randn(128, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
prompt = torch.randint(0, V - 1, (128,), device=dev, generator=g)
ks = bench.forced_schedule(seed=0)
forced = torch.tensor([ks], dtype=torch.int32, device=dev)
INJECT = os.environ.get("DFLASH_TRACE_NO_INJECT") is None
eng.reset_request(0, prompt, 1, 2048)
eng.prefill_context(0, ph)
if INJECT:
    eng.embed_block()


def step():
    if INJECT:
        eng._injected = "fresh"
        eng.draft_step_injected()
        eng.verify_step(tl, hsel, temperature=0.0, forced_k=forced, inject=True)
    else:
        eng._injected = "no"
        eng.draft_step()
        eng.verify_step(tl, hsel, temperature=0.0, forced_k=forced)


step()
torch.cuda.synchronize()
side = torch.cuda.Stream(device=dev)
graph = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    step()
torch.cuda.synchronize()
with torch.cuda.graph(graph, stream=side):
    step()
torch.cuda.synchronize()
for _ in range(60):
    graph.replay()
torch.cuda.synchronize()
cap = 4096
buf = torch.zeros(cap * 2, dtype=torch.int64, device=dev)
assert lib.dflash_step_trace_set(ctypes.c_void_p(buf.data_ptr()), cap) == 0
for _ in range(6):
    graph.replay()
torch.cuda.synchronize()
rec = buf.view(cap, 2).cpu()
rec = rec[rec[:, 1] != 0]
names = {(512, 32, 1): "finalize2", (256, 4, 16): "finalize_cl", (1024, 16, 1): "finalize_blk", (256, 192, 1): "qkv_post", (128, 16, 8): "attn_split",
         (256, 64, 1): "attn_combine", (256, 32, 16): "verify", (256, 18, 16): "verify", (192, 1, 148): "gemm", (192, 1, 132): "gemm_lm"}
ev = []
for tag, t in rec.tolist():
    phase, bd, gx, gy = tag & 15, (tag >> 4) & 0xFFF, (tag >> 16) & 0xFFFFFF, (tag >> 40) & 0xFFFFFF
    ev.append((t, names.get((bd, gx, gy), f"?{bd},{gx},{gy}"), phase))
ev.sort()
# one step = from a verify-kernel entry to the next
starts = [i for i, e in enumerate(ev) if e[1] == "verify" and e[2] == 0]
a, b = starts[-3], starts[-2]
t0 = ev[a][0]
ph_name = {0: "entry", 1: "past wait", 2: "end", 3: "first stage landed", 4: "last accumulator complete",
           5: "other CTAs' partials arrived"}
print(f"one step: {(ev[b][0] - t0) / 1e3:.1f} us, {b - a} records")
prev = t0
for t, n, p in ev[a:b]:
    print(f"{(t - t0) / 1e3:8.2f} us  (+{(t - prev) / 1e3:5.2f})  {n:13s} {ph_name[p]}")
    prev = t
