from .base import CNN, ContextGating  # noqa: F401
