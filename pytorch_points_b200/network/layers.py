"""The two layer wrappers the set-abstraction / feature-propagation modules need.

`Conv2d` and `SharedMLP` mirror `pytorch_points.network.layers.Conv2d` / `SharedMLP`
(network/layers.py:9-21,136-183): same constructor arguments, same sub-module names
(`conv`, `norm`, `act`, `layer{i}`) so state dicts interchange.  They are plain torch.nn
plumbing around the hot-path operators; the rest of the reference's layer zoo is out of scope
(SURVEY.md section 8)."""
from typing import List

import torch.nn as nn


class Conv2d(nn.Module):
    """2-D convolution followed by optional normalization and activation."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 activation=None, normalization=None, momentum=0.01, conv_params={}):
        super().__init__()
        self.activation = activation
        self.normalization = normalization
        bias = not normalization and bias
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              bias=bias, **conv_params)
        if normalization is not None:
            if normalization == "batch":
                self.norm = nn.BatchNorm2d(out_channels, affine=True, eps=0.001, momentum=momentum)
            elif normalization == "instance":
                self.norm = nn.InstanceNorm2d(out_channels, affine=True, eps=0.001, momentum=momentum)
            else:
                raise ValueError("only \"batch/instance\" normalization permitted.")
        if activation is not None:
            if activation == "relu":
                self.act = nn.ReLU()
            elif activation == "elu":
                self.act = nn.ELU(alpha=1.0)
            elif activation == "lrelu":
                self.act = nn.LeakyReLU(0.1)
            elif activation == "tanh":
                self.act = nn.Tanh()
            else:
                raise ValueError("only \"relu/elu/lrelu/tanh\" implemented")

    def forward(self, x, epoch=None):
        x = self.conv(x)
        if self.normalization is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.act(x)
        return x


class SharedMLP(nn.Sequential):
    """A stack of 1x1 `Conv2d` blocks applied to every (point, sample) position."""

    def __init__(self, args: List[int], activation: str = None, normalization: str = None, **kwargs):
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module("layer{}".format(i),
                            Conv2d(args[i], args[i + 1], 1, normalization=normalization, activation=activation))
