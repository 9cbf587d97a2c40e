"""Micro-benchmark of vgen_attention_d64 on the config-2 shapes (device time via CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vgen_b200 import ops  # noqa: E402


def main():
    g = torch.Generator().manual_seed(0)
    for (b, h, lq, lk, div) in [(16, 5, 14080, 14080, 1), (16, 10, 3520, 3520, 1), (16, 20, 880, 880, 1), (16, 5, 14080, 145, 16)]:
        inner = h * 64
        if lq == lk and div == 1:
            qkv = torch.randn(b, lq, 3 * inner, generator=g).half().cuda()
            q, k, v = qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:]
        else:
            q = torch.randn(b, lq, inner, generator=g).half().cuda()
            kv = torch.randn(b // div, lk, 2 * inner, generator=g).half().cuda()
            k, v = kv[:, :, :inner], kv[:, :, inner:]
        out = torch.empty(b, lq, inner, device="cuda", dtype=torch.float16)
        fn = lambda: ops.attention_d64(q, k, v, h, kv_batch_div=div, out=out)  # noqa: E731
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
        ms = ts[len(ts) // 2]
        flops = 4.0 * b * h * lq * lk * 64
        print({"shape": (b, h, lq, lk, div), "ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1)}, flush=True)


if __name__ == "__main__":
    main()
