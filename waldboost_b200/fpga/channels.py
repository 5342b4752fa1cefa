"""waldboost.fpga.channels on the GPU (reference waldboost/fpga/channels.py:29-66): integer gradient channels for uint8
images.  Like the functions of waldboost_b200.channels they are markers with a CUDA implementation: used as
`channel_opts["channels"]` they select the integer channel kernel, called on an image they run it on that image."""
import numpy as np

from .. import channels as _ch


def grad_hist_4_u1(image):
    """dx, (dx-dy)/2, dy, (dx+dy)/2 of the 3x3 Sobel stencils (border 0), min(|y| // 4, 255) -> (h, w, 4) uint8."""
    return _ch._direct(grad_hist_4_u1, _u8(image))


def grad_mag_u1(image):
    """min(max(|dx|, |dy|) // 4, 255) -> (h, w, 1) uint8."""
    return _ch._direct(grad_mag_u1, _u8(image))


def _u8(image):
    if not isinstance(image, np.ndarray) or image.dtype != np.uint8:
        raise TypeError("the FPGA integer channels take uint8 images")
    return image


_ch.register_channel_function(grad_hist_4_u1, dict(kind=_ch.N.WBG_CH_FPGA_HIST4_U1, n_bins=0, full=False, bias=0, norm=0, eps=0.0),
                              channels=4, integer=True)
_ch.register_channel_function(grad_mag_u1, dict(kind=_ch.N.WBG_CH_FPGA_MAG_U1, n_bins=0, full=False, bias=0, norm=0, eps=0.0),
                              channels=1, integer=True)
