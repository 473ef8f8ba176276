"""Table-driven construction of an nn.Module tree whose state_dict keys equal a given spec."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .synth import sinusoid_table


class Node(nn.Module):
    """Plain container; exists only so that dotted state_dict keys resolve like the reference's."""

    def extra_repr(self) -> str:
        own = [f"{k}{tuple(v.shape)}" for k, v in self._parameters.items()]
        return ", ".join(own)


def _leaf_parent(root: nn.Module, key: str):
    parts = key.split(".")
    mod = root
    for name in parts[:-1]:
        nxt = mod._modules.get(name)
        if nxt is None:
            nxt = Node()
            mod.add_module(name, nxt)
        mod = nxt
    return mod, parts[-1]


def _default_init(p, stats) -> torch.Tensor:
    """Same families as torch's defaults for the corresponding reference layers (random init == --restore_step 0)."""
    shape = tuple(p.shape)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = 1.0 / math.sqrt(max(fan_in, 1))
    if p.init in ("linear", "conv"):
        return torch.empty(shape).uniform_(-bound, bound)
    if p.init == "bias":
        return torch.empty(shape).uniform_(-0.05, 0.05)
    if p.init == "ln_w":
        return torch.ones(shape)
    if p.init in ("ln_b", "bn_mean"):
        return torch.zeros(shape)
    if p.init == "bn_var":
        return torch.ones(shape)
    if p.init == "embedding":
        return torch.randn(shape)
    if p.init == "embedding_pad0":
        w = torch.randn(shape)
        w[0].zero_()
        return w
    if p.init == "sinusoid":
        return sinusoid_table(shape[1], shape[2]).unsqueeze(0)
    if p.init == "pitch_bins":
        return torch.linspace(stats["pitch"][0], stats["pitch"][1], shape[0])
    if p.init == "energy_bins":
        return torch.linspace(stats["energy"][0], stats["energy"][1], shape[0])
    if p.init == "zero":
        return torch.zeros(shape, dtype=torch.long)
    if p.init == "wn_v":
        return torch.randn(shape) * 0.01
    if p.init == "wn_g":
        return torch.ones(shape)
    raise ValueError(p.init)


def populate(root: nn.Module, spec, stats=None):
    for p in spec:
        mod, leaf = _leaf_parent(root, p.key)
        val = _default_init(p, stats)
        if p.kind == "param":
            mod.register_parameter(leaf, nn.Parameter(val))
        elif p.kind == "frozen":
            mod.register_parameter(leaf, nn.Parameter(val, requires_grad=False))
        else:
            mod.register_buffer(leaf, val)


def get(root: nn.Module, key: str) -> torch.Tensor:
    mod = root
    parts = key.split(".")
    for name in parts[:-1]:
        mod = mod._modules[name]
    leaf = parts[-1]
    if leaf in mod._parameters:
        return mod._parameters[leaf]
    return mod._buffers[leaf]
