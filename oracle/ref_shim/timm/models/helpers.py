def load_pretrained(*args, **kwargs):
    raise RuntimeError("timm shim: load_pretrained is unavailable offline")
