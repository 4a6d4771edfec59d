from . import raw  # noqa: F401
from .sr_b200 import (channel_attention, conv_nhwc, conv_to_image, image_to_nhwc, nearest_up2, pad64, rcab, rcab_backward_floats,  # noqa: F401
                      res_block_nobn, shuffle_to_image)

__all__ = ['raw', 'channel_attention', 'conv_nhwc', 'conv_to_image', 'image_to_nhwc', 'nearest_up2', 'pad64', 'rcab', 'rcab_backward_floats', 'res_block_nobn',
           'shuffle_to_image']
