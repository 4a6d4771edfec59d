"""B200-native (sm_100a) super-resolution hot path for BasicSR4RS: EDSR / RCAN / SwinIR forward and
backward behind the reference's own ARCH_REGISTRY plugin surface.  See DESIGN.md."""
__version__ = '0.1.0'
