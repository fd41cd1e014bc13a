"""Checkpoint hand-off between the reference's training stages, for reference checkpoints loaded into the CUDA mirrors.

The reference saves ``net.state_dict()`` of a module that may be wrapped in ``nn.DataParallel`` (`src/utils/__init__.py:11-21`:
keys gain a ``module.`` prefix) and filters keys when one stage initialises the next:
  * MLM pre-training -> fine-tuning (`recipes/desed/finetune/passt/main.py:60-64`): drop ``classifier.`` and ``at_adpater.1`` keys;
  * PMAM post-pre-training -> later stages (`recipes/desed/pmam/main.py:188-191`): drop ``mlm_mlp.`` keys;
both with ``strict=False``.  The mirrors keep the reference's parameter names, so nothing is renamed."""
import torch

STAGE_FILTERS = {
    None: (),
    "finetune_from_mlm": ("classifier.", "at_adpater.1"),
    "pmam_from_existing": ("mlm_mlp.",),
}


def strip_data_parallel(state_dict):
    """Remove the ``module.`` prefix nn.DataParallel adds (only when every key carries it, as torch does)."""
    keys = list(state_dict)
    if keys and all(k.startswith("module.") for k in keys):
        return {k[len("module."):]: v for k, v in state_dict.items()}
    return dict(state_dict)


def filter_stage_keys(state_dict, stage):
    if stage not in STAGE_FILTERS:
        raise ValueError(f"unknown stage {stage!r}; choose from {sorted(k for k in STAGE_FILTERS if k)}")
    drop = STAGE_FILTERS[stage]
    return {k: v for k, v in state_dict.items() if not any(d in k for d in drop)}


def load_reference_checkpoint(model, checkpoint, stage=None, strict=None, map_location="cpu"):
    """Load a reference checkpoint (path or state dict) into a mirror module.  `stage` selects the upstream key filter; `strict`
    defaults to the reference's choice (False when a stage filter applies, True otherwise).  Returns torch's
    (missing_keys, unexpected_keys) result."""
    sd = torch.load(checkpoint, map_location=map_location) if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") else checkpoint
    if isinstance(sd, dict) and "state_dict" in sd and all(not torch.is_tensor(v) for v in sd.values()):
        sd = sd["state_dict"]
    sd = filter_stage_keys(strip_data_parallel(sd), stage)
    target = model.module if isinstance(model, torch.nn.DataParallel) else model
    if strict is None:
        strict = stage is None
    return target.load_state_dict(sd, strict=strict)
