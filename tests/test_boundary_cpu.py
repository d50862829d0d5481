"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports what include/dflash_b200.h
declares, the product class keeps the reference's state-dict / attribute contract, and compute entry points
fail loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from dflash_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from dflash_b200 import _lib
    names = _lib.exported_symbols()
    assert {"dflash_engine_create", "dflash_draft_step", "dflash_verify_step", "dflash_prefill_context",
            "dflash_sample", "dflash_gemm_skinny", "dflash_gemm_argmax", "dflash_workspace_bytes"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dflash_b200.h but not exported"
    assert lib.dflash_abi_version() == 3


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "dflash_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "at::" not in hdr and "#include <torch" not in hdr


def test_python_struct_matches_header():
    """ctypes mirror of dflash_config_t has the same field order as the header."""
    from dflash_b200.engine import BUFFERS, CConfig
    hdr = open(os.path.join(ROOT, "include", "dflash_b200.h")).read()
    body = re.search(r"typedef struct dflash_config \{(.*?)\} dflash_config_t;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|float|long long)\s+(\w+)\s*;", body)
    assert fields == [f[0] for f in CConfig._fields_]
    enum = re.search(r"enum dflash_buffer_id \{(.*?)\};", hdr, flags=re.S).group(1)
    enum = re.sub(r"/\*.*?\*/", "", enum, flags=re.S)
    ids = [x.strip().split("=")[0].strip() for x in enum.split(",") if x.strip()]
    assert ids[-1] == "DFLASH_BUF_COUNT" and len(ids) - 1 == len(BUFFERS)
    for cid, (name, _) in zip(ids, BUFFERS):
        assert cid == "DFLASH_BUF_" + name.upper(), (cid, name)


def test_stream_k_slot_math(lib):
    """dflash_gemm_max_slots is host code: cross-check the stream-K cut against a direct enumeration."""
    for N, K, grid in [(4096, 4096, 148), (6144, 4096, 148), (24576, 4096, 148), (4096, 12288, 148),
                       (4096, 20480, 148), (1000, 512, 37), (256, 128, 148), (4096, 4096, 7)]:
        nt, kb = (N + 127) // 128, K // 64
        T = nt * kb
        G = min(grid, T)
        owners = [set() for _ in range(nt)]
        covered = 0
        for g in range(G):
            u0, u1 = g * T // G, (g + 1) * T // G
            covered += u1 - u0
            for u in range(u0, u1):
                owners[u // kb].add(g)
        assert covered == T
        assert max(len(o) for o in owners) == lib.dflash_gemm_max_slots(N, K, grid)
        # slots of a tile are consecutive CTAs starting at the owner of its first unit
        for t, o in enumerate(owners):
            first = ((t * kb + 1) * G - 1) // T
            assert sorted(o) == list(range(first, first + len(o)))
    assert lib.dflash_gemm_max_slots(128, 100, 4) < 0  # K not a multiple of 64 -> error code


def test_bad_config_is_rejected(lib):
    from dflash_b200.engine import CConfig, _declare
    _declare(lib)
    cfg = CConfig(hidden=4096, intermediate=12288, n_layers=5, n_q_heads=32, n_kv_heads=8, head_dim=64, vocab=1000,
                  n_sel=5, block_size=16, max_requests=1, max_seq=4096, out_len=4096, hist_len=16, rms_eps=1e-6,
                  rope_scale=1.0, mask_token_id=1, gemm_grid=148)
    assert lib.dflash_workspace_bytes(ctypes.byref(cfg)) == 0
    assert b"head_dim" in lib.dflash_last_error()
    cfg.head_dim = 128
    n = lib.dflash_workspace_bytes(ctypes.byref(cfg))
    assert n > 5 * 2 * 8 * 4096 * 128 * 2  # at least the static draft KV cache
    cfg.max_requests = 3  # any stream count up to 64 (buffers are padded to the UMMA width)
    assert lib.dflash_workspace_bytes(ctypes.byref(cfg)) > n
    cfg.max_requests = 65
    assert lib.dflash_workspace_bytes(ctypes.byref(cfg)) == 0
    assert b"max_requests" in lib.dflash_last_error()


def test_product_class_keeps_reference_contract():
    from dflash_b200 import DFlashDraftModel
    from transformers.models.qwen3.modeling_qwen3 import Qwen3Config, Qwen3PreTrainedModel
    from tests.tiny_models import draft_config
    m = DFlashDraftModel(draft_config(16))
    assert isinstance(m, Qwen3PreTrainedModel) and DFlashDraftModel.config_class is Qwen3Config
    assert DFlashDraftModel._no_split_modules == ["Qwen3DFlashDecoderLayer"]
    assert m.block_size == 16 and m.mask_token_id == 999 and m.target_layer_ids == [1, 3]
    keys = set(m.state_dict().keys())
    expect = {"norm.weight", "fc.weight", "hidden_norm.weight"}
    for i in range(2):
        p = f"layers.{i}."
        expect |= {p + f"self_attn.{n}_proj.weight" for n in "qkvo"}
        expect |= {p + "self_attn.q_norm.weight", p + "self_attn.k_norm.weight"}
        expect |= {p + f"mlp.{n}_proj.weight" for n in ("gate", "up", "down")}
        expect |= {p + "input_layernorm.weight", p + "post_attention_layernorm.weight"}
    assert keys == expect  # rotary inv_freq is a non-persistent buffer
    assert m.fc.weight.shape == (256, 2 * 256)
    import inspect
    sig = inspect.signature(DFlashDraftModel.spec_generate)
    assert list(sig.parameters)[:6] == ["self", "target", "input_ids", "max_new_tokens", "stop_token_ids", "temperature"]
    sig = inspect.signature(DFlashDraftModel.forward)
    assert list(sig.parameters)[:7] == ["self", "position_ids", "attention_mask", "noise_embedding", "target_hidden",
                                        "past_key_values", "use_cache"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point raises instead of silently running elsewhere."""
    from dflash_b200 import DFlashDraftModel, DFlashNativeError, sample
    from tests.tiny_models import build_pair
    target, draft = build_pair(DFlashDraftModel, block_size=16, dtype=torch.bfloat16)
    prompt = torch.randint(0, 900, (1, 8))
    with pytest.raises(DFlashNativeError):
        draft.spec_generate(target, prompt, 4, None, 0.0)
    with pytest.raises(DFlashNativeError):
        sample(torch.zeros(1, 2, 16, dtype=torch.bfloat16), 0.0)


def test_product_never_imports_oracle():
    """The shipped package must not reach into oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "dflash_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_context_tap_equals_output_hidden_states():
    """forward hooks on the selected target layers give exactly hidden_states[id + 1]."""
    import torch
    from transformers import DynamicCache
    from dflash_b200 import ContextTap, build_target_layer_ids
    from tests.tiny_models import TINY, target_config, seeded_fill_
    from transformers import Qwen3ForCausalLM
    target = Qwen3ForCausalLM(target_config()).eval()
    seeded_fill_(target, 5)
    ids = build_target_layer_ids(TINY["target_layers"], TINY["draft_layers"])
    x = torch.randint(0, TINY["vocab"] - 1, (1, 19), generator=torch.Generator().manual_seed(2))
    with torch.inference_mode():
        ref = target(x, output_hidden_states=True, past_key_values=DynamicCache(), use_cache=True)
        n_hooks = [len(m._forward_hooks) for m in target.model.layers]  # transformers installs its own on first use
        with ContextTap(target, ids) as tap:
            out = target(x, past_key_values=DynamicCache(), use_cache=True)
    assert out.hidden_states is None
    assert torch.equal(out.logits, ref.logits)
    for s, i in zip(tap.states, ids):
        assert torch.equal(s, ref.hidden_states[i + 1])
    assert [len(m._forward_hooks) for m in target.model.layers] == n_hooks  # our hooks are removed on exit
    with pytest.raises(ValueError):
        ContextTap(target, [TINY["target_layers"] - 1])  # the last entry of hidden_states is post-norm, not a layer output


def test_whole_tile_grid_is_balanced(lib):
    """lm_head-style GEMMs run on the smallest grid with the same busiest CTA (host arithmetic, no GPU)."""
    lib.dflash_gemm_argmax_grid.restype = ctypes.c_int
    f = lib.dflash_gemm_argmax_grid
    assert f(151936, 148) == 132          # 1187 tiles: 9 per CTA either way, 132 CTAs all busy
    assert f(128256, 148) == 144          # 1002 tiles: 7 per CTA
    assert f(1000, 148) == 8              # fewer tiles than CTAs: one tile each
    assert f(151936, 37) == 36            # wide batches: 37 ranges -> 36 of 33 tiles
    for n_rows in (128, 129, 5000, 151936, 262144):
        for grid in (1, 7, 37, 148):
            g = f(n_rows, grid)
            nt = (n_rows + 127) // 128
            assert 1 <= g <= min(grid, nt)
            assert -(-nt // g) == -(-nt // min(grid, nt))   # same maximum tiles per CTA
    assert f(0, 148) < 0


def test_forced_tau_schedule_is_window_balanced():
    import bench
    ks = bench.forced_schedule(seed=0)
    mean = sum(k + 1 for k in ks) / len(ks)
    assert abs(mean - 7.28) < 0.01 and sorted(ks) == sorted(bench.forced_schedule(seed=0)) and max(ks) <= 15
    for w in (8, 16, 32):
        for i in range(0, len(ks) - w + 1, w):
            assert abs(sum(k + 1 for k in ks[i:i + w]) / w - mean) < 0.75


def test_automodel_from_pretrained_returns_the_product_class(tmp_path):
    """README.md:76-81: `AutoModel.from_pretrained(<draft>, trust_remote_code=True)`. A checkpoint directory with the
    installed remote-code module loads as this package's class, weights and DFlash attributes intact."""
    import torch
    from transformers import AutoModel
    from dflash_b200 import DFlashDraftModel
    from dflash_b200.hf import install_remote_code
    from tests.tiny_models import draft_config, seeded_fill_
    ref = DFlashDraftModel(draft_config(16))
    seeded_fill_(ref, 7)
    ref.save_pretrained(tmp_path)
    mod = install_remote_code(str(tmp_path))
    assert os.path.basename(mod) == "dflash.py"
    import json
    assert json.load(open(tmp_path / "config.json"))["auto_map"]["AutoModel"] == "dflash.DFlashDraftModel"
    m = AutoModel.from_pretrained(str(tmp_path), trust_remote_code=True)
    assert type(m).__name__ == "DFlashDraftModel" and isinstance(m, DFlashDraftModel)
    assert m.block_size == 16 and m.mask_token_id == ref.mask_token_id and m.target_layer_ids == ref.target_layer_ids
    sd_a, sd_b = ref.state_dict(), m.state_dict()
    assert sd_a.keys() == sd_b.keys()
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k
    install_remote_code(str(tmp_path))  # idempotent


def test_static_cache_is_a_transformers_cache_and_foreign_caches_track_length():
    import torch
    from transformers import DynamicCache
    from transformers.cache_utils import Cache
    from dflash_b200.model import DFlashStaticCache, _grow_foreign_cache
    c = DFlashStaticCache()
    assert isinstance(c, Cache) and c.get_seq_length() == 0
    c.length = 40
    c.crop(33)
    assert c.get_seq_length() == 33
    c.crop(100)
    assert c.get_seq_length() == 33
    with pytest.raises(RuntimeError):
        c.update(torch.zeros(1, 1, 1, 1), torch.zeros(1, 1, 1, 1), 0)
    d = DynamicCache()
    _grow_foreign_cache(d, 37, 2, "cpu")
    assert d.get_seq_length() == 37
    d.crop(21)
    _grow_foreign_cache(d, 5, 2, "cpu")
    assert d.get_seq_length() == 26


def test_shipped_library_has_no_result_changing_switches():
    """The product build reads no environment variable: timing experiments that change results live in debug builds."""
    import subprocess
    from dflash_b200 import _lib
    out = subprocess.run(["strings", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in ("DFLASH_DEBUG_SKIP", "DFLASH_LATE_W", "DFLASH_MEGA", "DFLASH_FUSED_ATTN", "DFLASH_NO_ROW_CLUSTER",
                 "DFLASH_NO_CARVEOUT", "DFLASH_LM_GRID"):
        assert name not in out, name
    src = "".join(open(os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", f)).read()
                  for f in os.listdir(os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc")))
    assert "getenv" not in src


def test_batched_ragged_target_forward_equals_per_request_forwards():
    """`BatchedVerifyTarget` (one verify forward for all streams: per-row positions, 4-D mask and a ragged static KV cache
    from start[r]) against the reference's per-request calls with a DynamicCache + crop (model/dflash.py:249-255,262),
    over three cycles with ragged acceptance lengths. Eager on CPU here; the GPU tests replay the same forward from a
    CUDA graph."""
    import torch
    from transformers import DynamicCache
    from dflash_b200 import DFlashDraftModel
    from dflash_b200.target_graph import BatchedVerifyTarget
    from dflash_b200.utils import ContextTap
    from tests.tiny_models import TINY, build_pair
    bs, R = 8, 3
    target, draft = build_pair(DFlashDraftModel, seed=1234, block_size=bs)
    layer_ids = draft.target_layer_ids
    g = torch.Generator().manual_seed(5)
    lens = [5, 21, 12]
    prompts = [torch.randint(0, TINY["vocab"] - 1, (1, n), generator=g) for n in lens]
    start = torch.tensor(lens, dtype=torch.int32)
    block_ids = torch.randint(0, TINY["vocab"] - 1, (R, bs), generator=g)
    bt = BatchedVerifyTarget(target, bs, R, 96, layer_ids, start, block_ids, bucket=32, use_graph=False)
    caches = []
    with torch.inference_mode():
        for r in range(R):
            logits0, hidden0 = bt.prefill(r, prompts[r])
            c = DynamicCache()
            with ContextTap(target, layer_ids) as tap:
                ref0 = target(prompts[r], position_ids=torch.arange(lens[r]).unsqueeze(0), past_key_values=c,
                              use_cache=True, logits_to_keep=1)
            torch.testing.assert_close(logits0, ref0.logits)
            for a, b in zip(hidden0, tap.states):
                torch.testing.assert_close(a, b)
            caches.append(c)
        for cyc, acc in enumerate([(3, 8, 1), (8, 1, 5), (2, 2, 2)]):
            max_end = int(start.max()) + bs
            logits, hidden = bt.verify_forward(max_end)
            assert bt.cache.kv_len == -(-max_end // 32) * 32 and logits.shape == (R * bs, TINY["vocab"])
            for r in range(R):
                s0 = int(start[r])
                with ContextTap(target, layer_ids) as tap:
                    ref = target(block_ids[r:r + 1], position_ids=torch.arange(s0, s0 + bs).unsqueeze(0),
                                 past_key_values=caches[r], use_cache=True)
                torch.testing.assert_close(logits[r * bs:(r + 1) * bs], ref.logits[0], rtol=1e-4, atol=1e-4)
                for a, b in zip(hidden, tap.states):
                    torch.testing.assert_close(a[r * bs:(r + 1) * bs], b[0], rtol=1e-4, atol=1e-4)
                caches[r].crop(s0 + acc[r])           # the reference's rollback ...
            start += torch.tensor(acc, dtype=torch.int32)  # ... is only the length here
            block_ids.copy_(torch.randint(0, TINY["vocab"] - 1, (R, bs), generator=g))
    assert bt.n_forwards == 3
