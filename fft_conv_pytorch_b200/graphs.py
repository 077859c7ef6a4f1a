"""CUDA-graph replay of a convolution call (launch-bound steady state: three kernels per call are cheaper to
replay as one graph than to queue one by one from Python)."""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor


class GraphedConv:
    """Captures ``fn(static_input)`` once and replays it.

    ``fn`` is any callable built from this package's ops (an ``FFTConv*`` module, a ``functools.partial`` of
    ``fft_conv`` ...) run under ``torch.no_grad()``. The captured graph owns its workspace and its output tensor:
    ``__call__`` returns the same output tensor every time (copy it if it must outlive the next call).
    Weights are read through the cached kernel spectrum captured at construction: re-capture after changing them.
    """

    def __init__(self, fn: Callable[[Tensor], Tensor], example_input: Tensor, warmup: int = 2):
        if not example_input.is_cuda:
            raise ValueError("GraphedConv needs a CUDA example input")
        self.static_input = example_input
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=example_input.device)
        side.wait_stream(torch.cuda.current_stream(example_input.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # builds plan, constant table and kernel spectrum outside the capture
                fn(self.static_input)
        torch.cuda.current_stream(example_input.device).wait_stream(side)
        torch.cuda.synchronize(example_input.device)
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_output = fn(self.static_input)

    def __call__(self, x: Tensor = None) -> Tensor:
        if x is not None and x.data_ptr() != self.static_input.data_ptr():
            self.static_input.copy_(x)
        self.graph.replay()
        return self.static_output


def graphed(fn: Callable[[Tensor], Tensor], example_input: Tensor) -> GraphedConv:
    return GraphedConv(fn, example_input)
