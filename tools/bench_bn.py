"""Short-K tap-GEMM: device time vs N tile (bn), with / without bias and residual (diagnosis of the epilogue-bound layers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vgen_b200 import ops  # noqa: E402


def timeit(fn, iters=8):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator().manual_seed(0)
for (m, k, n) in [(450560, 320, 960), (450560, 320, 320), (112640, 640, 1920)]:
    a = torch.randn(m, k, generator=g).half().cuda()
    w = (torch.randn(n, k, generator=g) * k ** -0.5).half().cuda()
    b = torch.randn(n, generator=g).cuda()
    r = torch.randn(m, n, generator=g).half().cuda()
    o = torch.empty(m, n, dtype=torch.float16, device="cuda")
    for bn in (0, 96, 128, 160, 192, 256):
        for (bias, res) in ((None, None), (b, None), (b, r)):
            ms = timeit(lambda: ops.linear(a, w, bias=bias, residual=res, out=o, bn=bn))
            byts = 2.0 * (m * k + m * n * (2 if res is not None else 1))
            print({"m": m, "k": k, "n": n, "bn": bn, "bias": bias is not None, "res": res is not None, "ms": round(ms, 4),
                   "tflops": round(2.0 * m * n * k / ms / 1e9, 1), "hbm_gbs": round(byts / ms / 1e6, 1)}, flush=True)
