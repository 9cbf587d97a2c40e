"""Run GroupNorm on the config-2 shapes a few times (for ncu) and print device timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vgen_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None
gn_cases = [(16, 14080, 640, True), (1, 225280, 320, True), (16, 14080, 320, True), (1, 56320, 640, True), (1, 14080, 1280, True), (16, 880, 1280, True)]
if only is not None:
    gn_cases = [] if only == "ln" else [gn_cases[int(only)]]
for (n, p, c, silu) in gn_cases:
    x = torch.randn(n, p, c, generator=g).half().cuda()
    ga, be = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    y = torch.empty_like(x)
    fn = lambda: ops.group_norm(x, ga, be, 1e-5, silu, out=y)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    ms = ts[2]
    print({"shape": (n, p, c), "ms": round(ms, 4), "GBs_rw": round(4.0 * x.numel() / ms / 1e6, 1), "GBs_3pass": round(6.0 * x.numel() / ms / 1e6, 1)}, flush=True)

for (rows, c) in ([(450560, 320), (112640, 640), (28160, 1280)] if only is None else ([(450560, 320)] if only == "ln" else [])):
    x = torch.randn(rows, c, generator=g).half().cuda()
    ga, be = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    y = torch.empty_like(x)
    fn = lambda: ops.layer_norm(x, ga, be, out=y)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print({"layer_norm": (rows, c), "ms": round(ts[2], 4), "GBs_rw": round(4.0 * x.numel() / ts[2] / 1e6, 1)}, flush=True)
