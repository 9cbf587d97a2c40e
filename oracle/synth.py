"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and inputs.

There are no checkpoints (no network), so both the oracle and the CUDA path run on synthetic
state_dicts.  Every tensor is generated from its NAME with numpy's PCG64 bit stream (stable across
platforms), so the GPU box regenerates exactly the weights the golden vectors were made with, and the
fixtures only need to store inputs' recipes and outputs.

Hygiene: the reference zero-initialises 158 tensors (every proj_out, the second
conv of each ResBlock, conv4 of each temporal block, the head conv, fps_embedding[-1]); with those at
zero the UNet output is a per-channel constant and any parity test is vacuous, so EVERY tensor gets
non-zero values here.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(name: str, seed: int):
    return np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), seed]))


def tensor(name: str, shape, scale: float = 1.0, seed: int = 0, dtype=torch.float32):
    a = _rng(name, seed).standard_normal(size=tuple(shape), dtype=np.float32) * np.float32(scale)
    return torch.from_numpy(a).to(dtype)


def state_dict(spec, seed: int = 0, gain: float = 1.0):
    """spec: iterable of (name, shape).  Weight matrices/filters ~ N(0, gain/sqrt(fan_in)),
    norm weights ~ 1 + 0.1 N, biases ~ 0.05 N."""
    sd = {}
    for name, shape in spec:
        shape = tuple(shape)
        is_norm = (".norm" in name or name.startswith("norm") or "in_layers.0." in name or
                   "out_layers.0." in name or name.startswith("out.0.") or
                   (".conv" in name and name.rsplit(".", 2)[-2] == "0" and "temopral_conv" in name))
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = tensor(name, shape, gain / np.sqrt(fan_in), seed)
        elif name.endswith("weight") and is_norm:
            sd[name] = 1.0 + tensor(name, shape, 0.1, seed)
        elif name.endswith("weight"):
            sd[name] = 1.0 + tensor(name, shape, 0.1, seed)
        else:
            sd[name] = tensor(name, shape, 0.05, seed)
    return sd


def spec_of(module):
    """(name, shape) list of a torch module's state_dict (used with the imported reference classes)."""
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]
