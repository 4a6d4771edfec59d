from . import raw  # noqa: F401
from .sr_b200 import (conv_nhwc, conv_to_image, image_to_nhwc, nearest_up2, pad64, rcab, res_block_nobn,  # noqa: F401
                      shuffle_to_image)

__all__ = ['raw', 'conv_nhwc', 'conv_to_image', 'image_to_nhwc', 'nearest_up2', 'pad64', 'rcab', 'res_block_nobn',
           'shuffle_to_image']
