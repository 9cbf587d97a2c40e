"""Micro-benchmark of individual tap-GEMM shapes (device time via CUDA events, L2 flushed between runs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vgen_b200 import ops

def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

def main():
    g = torch.Generator().manual_seed(0)
    cases = [("linear", 225280, 320, 320, True, False), ("linear", 225280, 320, 320, False, False), ("linear", 225280, 320, 960, False, False),
             ("linear", 225280, 320, 2560, False, True), ("linear", 225280, 1280, 320, True, False), ("linear", 56320, 640, 640, True, False),
             ("linear", 14080, 1280, 1280, True, False), ("conv", 16, 88, 160, 320, 320), ("conv", 16, 22, 40, 1280, 1280), ("conv", 16, 11, 20, 1280, 1280),
             ("tconv", 16, 14080, 320, 320)]
    out = []
    if len(sys.argv) > 2 and sys.argv[1] == "--only":       # ncu captures: one case, few iterations
        cases = [cases[int(sys.argv[2])]]
    for c in cases:
        if c[0] == "linear":
            _, m, k, n, res, geglu = c
            a = torch.randn(m, k, generator=g).half().cuda()
            w = (torch.randn(n, k, generator=g) * k ** -0.5).half().cuda()
            b = torch.randn(n, generator=g).cuda()
            r = torch.randn(m, n, generator=g).half().cuda() if res else None
            o = torch.empty(m, n // 2 if geglu else n, dtype=torch.float16, device="cuda")
            fn = lambda: ops.linear(a, w, bias=b, residual=r, geglu=geglu, out=o, bn=256 if geglu else 0)
            flops, byts = 2.0 * m * n * k, 2.0 * (m * k + m * o.shape[1] * (2 if res else 1))
        elif c[0] == "conv":
            _, nimg, h, w_, ci, n = c
            x = torch.randn(nimg, h, w_, ci, generator=g).half().cuda()
            w = (torch.randn(n, 9 * ci, generator=g) * (9 * ci) ** -0.5).half().cuda()
            b = torch.randn(n, generator=g).cuda()
            o = torch.empty(nimg, h, w_, n, dtype=torch.float16, device="cuda")
            fn = lambda: ops.conv2d_3x3(x, w, bias=b, out=o)
            flops, byts = 2.0 * nimg * h * w_ * n * 9 * ci, 2.0 * nimg * h * w_ * (ci + n)
        else:
            _, f, hw, ci, n = c
            x = torch.randn(f, hw, ci, generator=g).half().cuda()
            w = (torch.randn(n, 3 * ci, generator=g) * (3 * ci) ** -0.5).half().cuda()
            b = torch.randn(n, generator=g).cuda()
            o = torch.empty(f, hw, n, dtype=torch.float16, device="cuda")
            fn = lambda: ops.tconv3(x, w, bias=b, out=o)
            flops, byts = 2.0 * f * hw * n * 3 * ci, 2.0 * f * hw * (ci + n)
        ms = timeit(fn)
        rec = {"case": c, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "gbs": round(byts / ms / 1e6, 1)}
        out.append(rec)
        print(rec, flush=True)
    return out

if __name__ == "__main__":
    main()
