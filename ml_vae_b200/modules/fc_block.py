"""Drop-in for the reference's modules/fc_block.py:4-21.

Same constructor (fc_sizes, dropout=0.15 [ignored, as in the reference], end_activation)
and the same checkpoint keys (blocks.{0,2,4,...}.{weight,bias}); the forward runs
through ml_vae_b200.dense.linear_chain.
"""
from __future__ import annotations

from torch import nn

from ..dense import linear_chain
from ._params import attach, torch_default_linear


class FCBlock(nn.Module):
    def __init__(self, fc_sizes, dropout=0.15, end_activation=False):
        super().__init__()
        if len(fc_sizes) < 2:
            raise ValueError("fc_sizes needs at least an input and an output size")
        self.fc_sizes = [int(s) for s in fc_sizes]
        self.end_activation = bool(end_activation)
        self._w, self._b = [], []
        for i in range(len(self.fc_sizes) - 1):
            w, b = torch_default_linear(self.fc_sizes[i], self.fc_sizes[i + 1])
            # the reference interleaves LeakyReLU modules, so Linear layers sit at even indices
            self._w.append(attach(self, f"blocks.{2 * i}.weight", w))
            self._b.append(attach(self, f"blocks.{2 * i}.bias", b))

    def layers(self):
        ws = [self.blocks._modules[str(2 * i)].weight for i in range(len(self.fc_sizes) - 1)]
        bs = [self.blocks._modules[str(2 * i)].bias for i in range(len(self.fc_sizes) - 1)]
        return ws, bs

    def forward(self, x):
        ws, bs = self.layers()
        return linear_chain(x, ws, bs, self.end_activation)
