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


# (VGEN_ATTN_TILES token, VGEN_ATTN_STAGGER) variants of vgen_attention_d64 to compare; --variants a,b,c selects tokens
VARIANTS = [("auto", None), ("2", None), ("3", None), ("t1", None), ("t2", None)]


def set_variant(tok):
    """token = VGEN_ATTN_TILES value ('2', '3', 't1', 't2', 't3'), or 'auto' = the library's shape heuristic"""
    tok = str(tok)
    if tok == "auto":
        os.environ.pop("VGEN_ATTN_TILES", None)
    else:
        os.environ["VGEN_ATTN_TILES"] = tok


def sdpa_check(tokens):
    """Small parity cases (ragged lengths, shared context) of every selected variant against fp32 SDPA."""
    F = torch.nn.functional
    g = torch.Generator().manual_seed(5)
    worst = {}
    for (b, h, lq, lk, div) in [(1, 1, 128, 128, 1), (2, 2, 300, 300, 1), (2, 5, 880, 880, 1), (4, 5, 220, 77, 4), (4, 10, 3520, 145, 2),
                                (2, 2, 1000, 1, 1), (1, 2, 385, 64, 1), (1, 2, 400, 600, 1), (2, 1, 384, 65, 2)]:
        inner = h * 64
        q = torch.randn(b, lq, inner, generator=g).half().cuda()
        kv = torch.randn(b // div, lk, 2 * inner, generator=g).half().cuda()
        k, v = kv[:, :, :inner], kv[:, :, inner:]
        sp = lambda t: t.float().reshape(t.shape[0], t.shape[1], h, 64).permute(0, 2, 1, 3)  # noqa: E731
        ref = F.scaled_dot_product_attention(sp(q), sp(k.repeat_interleave(div, 0)), sp(v.repeat_interleave(div, 0)))
        ref = ref.permute(0, 2, 1, 3).reshape(b, lq, inner)
        for tok in tokens:
            set_variant(tok)
            out = torch.empty(b, lq, inner, device="cuda", dtype=torch.float16)
            call(q, k, v, out, h, div)
            torch.cuda.synchronize()
            err = float((out.float() - ref).abs().max() / ref.abs().max())
            worst[tok] = max(worst.get(tok, 0.0), err)
            if not err < 3e-3:
                print(json.dumps({"parity_fail": tok, "shape": (b, h, lq, lk, div), "err": err}), flush=True)
    print(json.dumps({"parity_worst_rel_err": worst}), flush=True)


def main():
    g = torch.Generator().manual_seed(0)
    res = []
    only = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--only" else None
    variants = VARIANTS
    if "--variants" in sys.argv:
        toks = sys.argv[sys.argv.index("--variants") + 1].split(",")
        variants = [(t, None) for t in toks]
        sdpa_check(toks)
    shapes = [(32, 5, 14080, 14080, 1), (32, 10, 3520, 3520, 1), (32, 20, 880, 880, 1), (32, 20, 220, 220, 1),
              (32, 5, 14080, 145, 16), (32, 10, 3520, 145, 16)]
    if only is not None:
        shapes = [] if only == "d512" else [shapes[int(x)] for x in only.split(",")]
    for (b, h, lq, lk, div) in shapes:
        inner = h * 64
        if lq == lk and div == 1:
            qkv = torch.randn(b, lq, 3 * inner, generator=g).half().cuda()
            q, k, v = qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:]
        else:
            q = torch.randn(b, lq, inner, generator=g).half().cuda()
            kv = torch.randn(b // div, lk, 2 * inner, generator=g).half().cuda()
            k, v = kv[:, :, :inner], kv[:, :, inner:]
        ref_out = None
        for (tiles, stagger) in variants:
            set_variant(tiles)
            if stagger is None:
                os.environ.pop("VGEN_ATTN_STAGGER", None)
            else:
                os.environ["VGEN_ATTN_STAGGER"] = str(stagger)
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
            if ref_out is None:
                ref_out = out
            dmax = float((out.float() - ref_out.float()).abs().max())
            row = {"shape": (b, h, lq, lk, div), "tiles": tiles, "stagger": stagger, "maxdiff_vs_first": dmax, "ms": round(ms, 4), "min_ms": round(ts[0], 4),
                   "tflops": round(flops / ms / 1e9, 1),
                   "phase_cycles_per_block": {f"tile{i}": {n: round(t[i * 8 + j] / max(t[i * 8 + 5], 1), 1)
                                                          for j, n in enumerate(["wait_S", "ld_max", "wait_Pbuf", "exp_store", "total"])}
                                              for i in (0, 1)}}
            print(json.dumps(row), flush=True)
            res.append(row)
    os.environ.pop("VGEN_ATTN_TILES", None)
    os.environ.pop("VGEN_ATTN_STAGGER", None)
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
