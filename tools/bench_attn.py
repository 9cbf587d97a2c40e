"""Micro-benchmark of vgen_attention_d64 / vgen_attention_d512 on the config-2 shapes (device time via CUDA events), plus
the instrumented twin's per-phase cycle counters (softmax warp 0 of each q-tile of one mid-grid CTA)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vgen_b200 import lib  # noqa: E402


def call(q, k, v, out, h, div, timing=None):
    L = lib.load()
    b, lq, inner = q.shape
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if timing is None:
        rc = L.vgen_attention_d64(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, h, lq, k.shape[1], q.stride(1),
                                  k.stride(1), v.stride(1), out.stride(1), div, 64 ** -0.5, st)
    else:
        rc = L.vgen_attention_d64_debug(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, h, lq, k.shape[1], q.stride(1),
                                        k.stride(1), v.stride(1), out.stride(1), div, 64 ** -0.5, timing.data_ptr(), st)
    lib.check(rc, "vgen_attention_d64")


def main():
    g = torch.Generator().manual_seed(0)
    res = []
    only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None
    shapes = [(32, 5, 14080, 14080, 1), (32, 10, 3520, 3520, 1), (32, 20, 880, 880, 1), (32, 20, 220, 220, 1),
              (32, 5, 14080, 145, 16), (32, 10, 3520, 145, 16)]
    if only is not None:
        shapes = [] if only == "d512" else [shapes[int(only)]]
    for (b, h, lq, lk, div) in shapes:
        inner = h * 64
        if lq == lk and div == 1:
            qkv = torch.randn(b, lq, 3 * inner, generator=g).half().cuda()
            q, k, v = qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:]
        else:
            q = torch.randn(b, lq, inner, generator=g).half().cuda()
            kv = torch.randn(b // div, lk, 2 * inner, generator=g).half().cuda()
            k, v = kv[:, :, :inner], kv[:, :, inner:]
        for _once in (0,):
            out = torch.empty(b, lq, inner, device="cuda", dtype=torch.float16)
            fn = lambda: call(q, k, v, out, h, div)  # noqa: E731
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts.sort()
            ms = ts[len(ts) // 2]
            flops = 4.0 * b * h * lq * lk * 64
            timing = torch.zeros(16, dtype=torch.int64, device="cuda")
            call(q, k, v, out, h, div, timing)
            torch.cuda.synchronize()
            t = timing.cpu().tolist()
            row = {"shape": (b, h, lq, lk, div), "ms": round(ms, 4), "min_ms": round(ts[0], 4),
                   "tflops": round(flops / ms / 1e9, 1),
                   "phase_cycles_per_block": {f"tile{i}": {n: round(t[i * 8 + j] / max(t[i * 8 + 5], 1), 1)
                                                          for j, n in enumerate(["wait_S", "ld_max", "wait_Pbuf", "exp_store", "total"])}
                                              for i in (0, 1)}}
            print(json.dumps(row), flush=True)
            res.append(row)
    from vgen_b200 import ops
    for (b, l) in ([(2, 14080), (4, 14400), (2, 1792)] if only in (None, "d512") else []):
        qkv = torch.randn(b, l, 1536, generator=g).half().cuda()
        q, k, v = qkv[:, :, :512], qkv[:, :, 512:1024], qkv[:, :, 1024:]
        out = torch.empty(b, l, 512, device="cuda", dtype=torch.float16)
        ops.attention_d512(q, k, v, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.attention_d512(q, k, v, out=out)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        row = {"d512_shape": (b, l), "ms": round(ts[2], 4), "tflops": round(4.0 * b * l * l * 512 / ts[2] / 1e9, 1)}
        print(json.dumps(row), flush=True)
        res.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_attn.json", "w"), indent=0)


if __name__ == "__main__":
    main()
