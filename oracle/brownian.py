"""TEST INFRASTRUCTURE ONLY -- stand-in for torchsde.BrownianTree (torchsde is a third-party dependency
of the reference, requirements.txt `torchsde`, not installed in this image and not vendored).

tools/modules/diffusions/diffusion_gauss.py:22-49 uses exactly this interface:
    tree = torchsde.BrownianTree(t0, w0, t1, entropy=seed)      # w0: zeros_like(x)
    w = tree(ta, tb)                                            # -> W(tb) - W(ta), shape of w0
The increments must be consistent (W(a,c) = W(a,b) + W(b,c)) and N(0, |tb-ta|) distributed; torchsde
gets that from a lazily-refined Brownian-bridge tree seeded by `entropy`.  The random STREAM of the real
library cannot be reproduced without it, so parity of the SDE sampler is checked with this stub plugged
into BOTH the reference and the product (same seed -> same increments); the deterministic part of the
sampler (sigma schedule, exponential-integrator coefficients, 2M correction) is pinned exactly.

Construction: W is sampled on demand at each queried time by Brownian-bridge interpolation between the
nearest already-known times (a sorted cache), with a per-time generator seed derived from (entropy,
time); the path is reproducible for a given (entropy, query order) -- the reference and the oracle
query in the same order.
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
        self.shape, self.dtype, self.device = w0.shape, w0.dtype, w0.device
        self.entropy = int(entropy) if entropy is not None else 0
        self._times = [self.t0, self.t1]
        self._vals = [torch.zeros(self.shape, dtype=torch.float64),
                      self._normal(self.t1) * (self.t1 - self.t0) ** 0.5]

    def _normal(self, t):
        key = struct.unpack("<q", struct.pack("<d", float(t)))[0]
        g = torch.Generator().manual_seed(_mix64(_mix64(self.entropy) ^ key))
        return torch.randn(self.shape, generator=g, dtype=torch.float64)

    def _w(self, t):
        t = min(max(float(t), self.t0), self.t1)
        i = bisect.bisect_left(self._times, t)
        if i < len(self._times) and self._times[i] == t:
            return self._vals[i]
        ta, tb = self._times[i - 1], self._times[i]
        wa, wb = self._vals[i - 1], self._vals[i]
        mean = wa + (wb - wa) * ((t - ta) / (tb - ta))
        std = ((tb - t) * (t - ta) / (tb - ta)) ** 0.5
        w = mean + std * self._normal(t)
        self._times.insert(i, t)
        self._vals.insert(i, w)
        return w

    def __call__(self, ta, tb):
        return (self._w(tb) - self._w(ta)).to(device=self.device, dtype=self.dtype)
