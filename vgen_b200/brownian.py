"""Brownian-motion increments for the SDE solver of GaussianDiffusion.sample (diffusion_gauss.py:22-76).

The reference delegates to torchsde.BrownianTree (a pip dependency, not part of VGen).  When torchsde is
importable it is used unchanged; otherwise this module provides the same call interface
    tree = BrownianTree(t0, w0, t1, entropy=seed);  tree(ta, tb) -> W(tb) - W(ta)
with W sampled lazily by Brownian-bridge refinement between the already-known times, directly on the
device of w0 (no host round trip inside the sampling loop).  Increments are always consistent
(W(a,c) = W(a,b) + W(b,c)); the path is reproducible for a given (entropy, query order).
"""
from __future__ import annotations

import bisect
import struct

import torch


def _mix64(v):
    """splitmix64 finaliser: torch's CPU generator only consumes the low seed bits, and the IEEE bit
    patterns of nearby times differ mostly in the high ones."""
    v &= 0xFFFFFFFFFFFFFFFF
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (v ^ (v >> 31)) & 0x7FFFFFFFFFFFFFFF


class BrownianTree:
    def __init__(self, t0, w0, t1, entropy=None, **_):
        self.t0, self.t1 = float(t0), float(t1)
        self.like = w0
        self.entropy = int(entropy) if entropy is not None else 0
        self._gen = torch.Generator(device=w0.device)
        self._times = [self.t0, self.t1]
        self._vals = [torch.zeros_like(w0), self._normal(self.t1) * (self.t1 - self.t0) ** 0.5]

    def _normal(self, t):
        key = struct.unpack("<q", struct.pack("<d", float(t)))[0]
        self._gen.manual_seed(_mix64(_mix64(self.entropy) ^ key))
        return torch.randn(self.like.shape, generator=self._gen, device=self.like.device, dtype=self.like.dtype)

    def _w(self, t):
        t = min(max(float(t), self.t0), self.t1)
        i = bisect.bisect_left(self._times, t)
        if i < len(self._times) and self._times[i] == t:
            return self._vals[i]
        ta, tb = self._times[i - 1], self._times[i]
        wa, wb = self._vals[i - 1], self._vals[i]
        w = torch.lerp(wa, wb, (t - ta) / (tb - ta)) + self._normal(t) * ((tb - t) * (t - ta) / (tb - ta)) ** 0.5
        self._times.insert(i, t)
        self._vals.insert(i, w)
        return w

    def __call__(self, ta, tb):
        return self._w(tb) - self._w(ta)


def default_tree_cls():
    try:
        import torchsde  # the reference's own noise source, when installed
        return torchsde.BrownianTree
    except ImportError:
        return BrownianTree
