"""`AutoModel.from_pretrained(<draft checkpoint>, trust_remote_code=True)` route (reference README.md:76-81).

The published DFlash checkpoints carry `"auto_map": {"AutoModel": "dflash.DFlashDraftModel"}` in their config.json and
a `dflash.py` next to it (the reference's model/dflash.py); transformers imports that file and instantiates the class
it names. `install_remote_code(checkpoint_dir)` puts this package's class behind the same two names: it writes a
`dflash.py` that re-exports `dflash_b200.DFlashDraftModel` and makes sure config.json carries the `auto_map` entry.
Nothing else about the checkpoint changes (same weights file, same state-dict keys), so the user's loading line stays

    model = AutoModel.from_pretrained(path, trust_remote_code=True, dtype=torch.bfloat16).to("cuda")
"""
from __future__ import annotations

import json
import os

AUTO_MAP = {"AutoModel": "dflash.DFlashDraftModel"}

REMOTE_MODULE = '''"""Remote-code entry of a DFlash draft checkpoint: the B200-native implementation.

Same class contract as the reference's model/dflash.py (a Qwen3PreTrainedModel with the same sub-module names,
`spec_generate`, `forward`, `block_size`, `mask_token_id`, `target_layer_ids`); the arithmetic runs in
libdflash_b200.so (sm_100a). Requires the `dflash_b200` package on sys.path.
"""
from dflash_b200.model import (DFlashDraftModel, DFlashStaticCache, Qwen3DFlashAttention,  # noqa: F401
                               Qwen3DFlashDecoderLayer)
from dflash_b200.utils import build_target_layer_ids, extract_context_feature, sample  # noqa: F401
'''


def install_remote_code(checkpoint_dir: str, module_name: str = "dflash") -> str:
    """Write `<checkpoint_dir>/<module_name>.py` (a re-export of this package's classes) and add the `auto_map` entry
    to `<checkpoint_dir>/config.json`. Returns the path of the module file. Idempotent."""
    cfg_path = os.path.join(checkpoint_dir, "config.json")
    if not os.path.isfile(cfg_path):
        raise FileNotFoundError(f"{cfg_path}: not a checkpoint directory")
    mod_path = os.path.join(checkpoint_dir, module_name + ".py")
    with open(mod_path, "w") as f:
        f.write(REMOTE_MODULE)
    with open(cfg_path) as f:
        cfg = json.load(f)
    auto_map = dict(cfg.get("auto_map") or {})
    auto_map["AutoModel"] = f"{module_name}.DFlashDraftModel"
    cfg["auto_map"] = auto_map
    cfg.setdefault("architectures", ["DFlashDraftModel"])
    with open(cfg_path, "w") as f:
        json.dump(cfg, f, indent=2)
    return mod_path
