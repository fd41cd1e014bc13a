from .layers import Linear, LoRALayer, lora_state_dict, mark_only_lora_as_trainable  # noqa: F401
