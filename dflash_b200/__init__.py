"""dflash_b200 — B200-native DFlash draft-and-verify hot path behind the reference's Python API.

    from dflash_b200 import DFlashDraftModel            # == reference `from model import DFlashDraftModel`
    draft = DFlashDraftModel.from_pretrained(path, dtype=torch.bfloat16).to("cuda")
    out = draft.spec_generate(target, input_ids, max_new_tokens, stop_token_ids, temperature)

The arithmetic runs in hand-written sm_100a CUDA (libdflash_b200.so, C ABI in include/dflash_b200.h).
There is no CPU or PyTorch fallback: without the library or an sm_100 GPU the compute calls raise.
"""
from ._lib import DFlashNativeError, LIB_PATH  # noqa: F401


def __getattr__(name):
    # heavy imports (torch/transformers) only when the model classes are actually requested
    if name in ("DFlashDraftModel", "DFlashStaticCache", "Qwen3DFlashDecoderLayer", "Qwen3DFlashAttention"):
        from . import model
        return getattr(model, name)
    if name in ("extract_context_feature", "sample", "build_target_layer_ids", "select_context_states",
                "ContextTap"):
        from . import utils
        return getattr(utils, name)
    if name in ("dflash_generate", "dflash_generate_candidates", "cuda_time"):
        from . import generate
        return getattr(generate, name)
    if name in ("EwmaBlockScheduler",):
        from . import schedule
        return getattr(schedule, name)
    if name in ("spec_generate_batch",):
        from . import batched
        return getattr(batched, name)
    if name in ("DraftEngine",):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
