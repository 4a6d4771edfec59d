"""Arch plugin surface: ``build_network(opt)`` looks ``opt['type']`` up in ``ARCH_REGISTRY`` and calls it with
the remaining YAML keys, exactly like the reference's ``basicsr/archs/__init__.py`` (:17-23)."""
from copy import deepcopy

from ..utils.registry import ARCH_REGISTRY
from . import arch_util, edsr_arch, rcan_arch, swinir_arch  # noqa: F401  (importing registers the archs)

__all__ = ['build_network', 'ARCH_REGISTRY']


def build_network(opt):
    opt = deepcopy(opt)
    network_type = opt.pop('type')
    return ARCH_REGISTRY.get(network_type)(**opt)
