"""``nn.Module`` API: drop-in ``FFTConv{1,2,3}d`` / ``FFTConvTranspose{1,2,3}d``.

Same construction as the reference (reference fft_conv_pytorch/nn.py:42-63): each class inherits the constructor,
parameters (``weight``, ``bias``) and ``state_dict`` layout of ``torch.nn.Conv{n}d`` / ``ConvTranspose{n}d`` and only
replaces ``forward`` (reference nn.py:10-22, 28-39), so checkpoints interchange with the torch layers.
The kernel spectrum is cached per weight version inside ``functional`` and is never part of ``state_dict``.
"""
from torch import Tensor, nn

from .functional import fft_conv, fft_conv_transpose


class _FFTConvForward(nn.Module):
    """Forward of the FFT convolution layers (reference nn.py:7-22)."""

    def forward(self, signal: Tensor) -> Tensor:
        assert signal.ndim == self.weight.ndim  # batched input only, like the reference (nn.py:11)
        padding_mode = "constant" if self.padding_mode == "zeros" else self.padding_mode  # nn.py:12
        return fft_conv(
            signal,
            self.weight,
            bias=self.bias,
            stride=self.stride,
            padding=self.padding,
            dilation=self.dilation,
            groups=self.groups,
            padding_mode=padding_mode,
        )


class _FFTConvTransposeForward(nn.Module):
    """Forward of the transposed FFT convolution layers (reference nn.py:25-39; no ``output_size`` argument)."""

    def forward(self, signal: Tensor) -> Tensor:
        assert signal.ndim == self.weight.ndim
        return fft_conv_transpose(
            signal,
            self.weight,
            bias=self.bias,
            stride=self.stride,
            padding=self.padding,
            output_padding=self.output_padding,
            dilation=self.dilation,
            groups=self.groups,
        )


class FFTConv1d(_FFTConvForward, nn.Conv1d):
    ...


class FFTConv2d(_FFTConvForward, nn.Conv2d):
    ...


class FFTConv3d(_FFTConvForward, nn.Conv3d):
    ...


class FFTConvTranspose1d(_FFTConvTransposeForward, nn.ConvTranspose1d):
    ...


class FFTConvTranspose2d(_FFTConvTransposeForward, nn.ConvTranspose2d):
    ...


class FFTConvTranspose3d(_FFTConvTransposeForward, nn.ConvTranspose3d):
    ...
