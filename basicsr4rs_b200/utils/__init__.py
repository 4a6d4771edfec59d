from .registry import ARCH_REGISTRY, Registry  # noqa: F401
