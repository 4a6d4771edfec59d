from . import raw  # noqa: F401
from .sr_b200 import conv_nhwc, conv_to_image, image_to_nhwc, pad64, rcab, res_block_nobn  # noqa: F401

__all__ = ['raw', 'conv_nhwc', 'conv_to_image', 'image_to_nhwc', 'pad64', 'rcab', 'res_block_nobn']
