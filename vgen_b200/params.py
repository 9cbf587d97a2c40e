"""Parameter container: an nn.Module tree whose state_dict keys are exactly the reference's dotted
names, built from a (name, shape) spec.  The classes that own the CUDA forward derive from it so that
`load_state_dict(strict=True)`, `.to(gpu)`, `.eval()`, `.parameters()` and DistributedDataParallel
wrapping behave like the reference modules."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class _Node(nn.Module):
    """Anonymous container; children and parameters are added by path component."""


class SpecModule(nn.Module):
    def __init__(self):
        super().__init__()
        self._packed = None  # device-side packed weights, rebuilt lazily

    def _build_params(self, spec, zero_init=None, seed=0):
        """Create fp32 parameters for every (name, shape).  Default initialisation mirrors the spirit
        of the reference's (fan-in scaled weights, unit norms, the reference's zero-initialised
        tensors at zero); real runs load a checkpoint over it."""
        g = torch.Generator().manual_seed(seed)
        for name, shape in spec:
            parts = name.split(".")
            node = self
            for comp in parts[:-1]:
                child = node._modules.get(comp)
                if child is None:
                    child = _Node()
                    node.add_module(comp, child)
                node = child
            if zero_init is not None and zero_init(name):
                t = torch.zeros(shape)
            elif len(shape) >= 2:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
            elif parts[-1] == "weight":
                t = torch.ones(shape)
            else:
                t = torch.zeros(shape)
            node.register_parameter(parts[-1], nn.Parameter(t))

    # ---- packed-weight cache invalidation -------------------------------------------------------
    def invalidate_packed(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._masters_offloaded = False
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    @property
    def device(self):
        if self._masters_offloaded and self._exec_device is not None:
            return self._exec_device
        return next(self.parameters()).device

    _masters_offloaded = False
    _exec_device = None

    def packed_tensors(self):
        """(name, tensor) of the device-side packed weights (fp16 GEMM matrices, fp32 bias / affine vectors) in a
        deterministic order -- what the forward actually reads; packs on first use."""
        W = self._packed or self._pack()
        return [(k, W[k]) for k in sorted(W) if torch.is_tensor(W[k])]

    def offload_masters(self):
        """Inference-only memory diet: keep just the packed fp16 arena on the device and move the fp32 master
        parameters (reference key names; needed again only for state_dict() / re-packing) to host memory.
        The packed weights and captured graphs stay valid; a later .to(device) / load_state_dict() re-packs."""
        W = self._packed or self._pack()
        self._exec_device = next(self.parameters()).device
        for p in self.parameters():
            p.data = p.data.to("cpu")
        self._packed = W                      # .data assignment does not go through _apply: nothing was invalidated
        self._masters_offloaded = True
        return self
