"""Parameter plumbing shared by the drop-in modules.

The reference builds its modules out of nested nn.Sequential / FCBlock objects,
which fixes the checkpoint key names (e.g. ``fc.0.blocks.2.weight``).  Here the
modules are flat kernels-over-weights objects; ``Scope`` containers exist only so
that ``state_dict()`` / ``load_state_dict()`` expose exactly the reference's keys
(Checkpointer compatibility).
"""
from __future__ import annotations

import torch
from torch import nn


class Scope(nn.Module):
    """Name-space node: holds parameters/sub-scopes, has no forward."""


def attach(root: nn.Module, dotted: str, param: nn.Parameter) -> nn.Parameter:
    """Register ``param`` on ``root`` under a dotted reference key such as 'fc.0.blocks.2.weight'."""
    *scopes, leaf = dotted.split(".")
    node = root
    for s in scopes:
        nxt = node._modules.get(s)
        if nxt is None:
            nxt = Scope()
            node.add_module(s, nxt)
        node = nxt
    node.register_parameter(leaf, param)
    return param


def torch_default_linear(fan_in: int, fan_out: int):
    """(weight, bias) drawn exactly like nn.Linear(fan_in, fan_out) would, consuming the
    global RNG identically, so seeding like the reference (run.yaml:2-3) gives its weights."""
    lin = nn.Linear(fan_in, fan_out)
    return nn.Parameter(lin.weight.detach()), nn.Parameter(lin.bias.detach())


def torch_default_lstm(input_size: int, hidden: int, layers: int):
    """Parameters of nn.LSTM(input, hidden, layers, bidirectional=True) in torch's own order and init."""
    ref = nn.LSTM(input_size, hidden, layers, bidirectional=True, batch_first=True)
    return [(n, nn.Parameter(p.detach())) for n, p in ref.named_parameters()]
