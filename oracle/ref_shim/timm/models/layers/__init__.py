from torch.nn.init import trunc_normal_  # noqa: F401

from .helpers import to_2tuple  # noqa: F401
