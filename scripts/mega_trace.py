"""Per-phase timeline of the persistent step kernel (CTA 0's view): DFLASH_MEGA=1 DFLASH_MEGA_TRACE=1."""
import os, sys, torch
os.environ["DFLASH_MEGA"] = "1"; os.environ["DFLASH_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Q8, PROMPT_LEN, MAX_NEW, build_engine
dev = torch.device("cuda:0")
draft, eng, embed, lm_head = build_engine(Q8, dev, 0)
g = torch.Generator(device=dev).manual_seed(1)
H, L = Q8["hidden"], Q8["draft_layers"]
ph = [(torch.randn(PROMPT_LEN, H, device=dev, generator=g) * 0.5).to(torch.bfloat16) for _ in range(L)]
eng.reset_request(0, torch.randint(0, 1000, (PROMPT_LEN,), device=dev), 1, MAX_NEW)
eng.prefill_context(0, ph)
for _ in range(5):
    eng.draft_step()
torch.cuda.synchronize()
sync = eng.buf["mega_sync"].cpu()
dup = "DFLASH_MEGA_DUP" in os.environ
nph = 3 + (12 if dup else 10) * L + 2
base = 8 + 3 + 12 * L + 2
tr = sync[base: base + 3 * nph + 1].tolist()
t0 = tr[nph]
kinds = ["rows(embed)", "GEMM fc", "rows(fc)"]
for l in range(L):
    kinds += ["GEMM qkv", "qkv_post", "attn", "combine", "GEMM o", "rows(o)"] + (["rows(o)#2"] if dup else []) + ["GEMM gu", "swiglu", "GEMM d", "rows(d)"] + (["rows(d)#2"] if dup else [])
kinds += ["GEMM lm", "tokens"]
prev = t0
print("epoch", int(sync[0]), "err", int(sync[1]) & 0xffffffff)
for p in range(nph):
    w, d = tr[nph + 1 + p], tr[2 * nph + 1 + p]
    extra = f" wait {(w-prev)/1000:6.2f} work {(d-w)/1000:6.2f} arrive {(tr[p]-d)/1000:5.2f}" if not kinds[p].startswith("GEMM") else f" (own share done +{(d-prev)/1000:6.2f})"
    print(f"{p:3d} {kinds[p]:12s} +{(tr[p]-prev)/1000:8.2f} us   (t={(tr[p]-t0)/1000:8.1f}){extra}")
    prev = tr[p]
