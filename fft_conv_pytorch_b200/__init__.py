"""B200-native drop-in for the convolution path of fft-conv-pytorch.

Exports the reference's public names (reference fft_conv_pytorch/__init__.py:1-9) plus the functional entry points
(the reference's README imports ``fft_conv`` from the package root, which its ``__init__`` forgets to export).
See DESIGN.md for the architecture and INTEGRATION.md for the C ABI.
"""
from . import functional, nn
from .functional import complex_matmul, fft_conv, fft_conv_transpose
from .graphs import GraphedConv, graphed
from .nn import (
    FFTConv1d,
    FFTConv2d,
    FFTConv3d,
    FFTConvTranspose1d,
    FFTConvTranspose2d,
    FFTConvTranspose3d,
)
from .utils import to_ntuple

__version__ = "0.1.0"
__all__ = [
    "functional", "nn", "fft_conv", "fft_conv_transpose", "complex_matmul", "to_ntuple", "graphed", "GraphedConv",
    "FFTConv1d", "FFTConv2d", "FFTConv3d", "FFTConvTranspose1d", "FFTConvTranspose2d", "FFTConvTranspose3d",
]
