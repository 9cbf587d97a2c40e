"""GPU bring-up checks: run each op of libvgen_b200.so against a plain PyTorch fp32 reference and
print one line per case (never stops at the first failure, so one gpurun call yields a full picture).

  python tools/gpu_check.py --group tapgemm          # one group in-process
  python tools/gpu_check.py --all                    # every group, each in its own subprocess + timeout

This is a diagnostic tool; the pytest -m gpu suite is the gate.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

RESULTS = []


def rel_err(a, b):
    a = a.float()
    b = b.float()
    denom = b.abs().max().item() + 1e-12
    return (a - b).abs().max().item() / denom


def report(name, err, tol, extra=""):
    ok = err == err and err <= tol
    RESULTS.append({"name": name, "err": err, "tol": tol, "ok": bool(ok)})
    print(f"{'PASS' if ok else 'FAIL'}  {name:58s} err={err:.3e} tol={tol:.1e} {extra}", flush=True)


def run_case(name, fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        RESULTS.append({"name": name, "err": float("nan"), "ok": False, "exc": repr(e)})
        print(f"FAIL  {name:58s} EXC {e!r}", flush=True)
        traceback.print_exc()


# ----------------------------------------------------------------------------------------------
def group_tapgemm(impl="sm100"):
    from vgen_b200 import ops
    ops.set_tapgemm_impl(impl)
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).to(dev)

    def lin_case(m, k, n, bias=False, res=False, alpha=1.0, bn=0, lda_pad=0):
        def f():
            a_full = rnd(m, k + lda_pad).half()
            a = a_full[:, :k]
            w = rnd(n, k, scale=k ** -0.5).half()
            b = rnd(n) if bias else None
            r = rnd(m, n).half() if res else None
            out = ops.linear(a, w, bias=b, residual=r, alpha=alpha, bn=bn)
            torch.cuda.synchronize()
            ref = (a.float() @ w.float().t()) * alpha
            if b is not None:
                ref = ref + b
            if r is not None:
                ref = ref.half().float() + r.float()
            report(f"[{impl}] linear m{m} k{k} n{n} bias{int(bias)} res{int(res)} a{alpha} bn{bn} pad{lda_pad}",
                   rel_err(out, ref), 2e-3)
        run_case(f"linear m{m} k{k} n{n}", f)

    lin_case(128, 64, 32)
    lin_case(128, 64, 128)
    lin_case(128, 128, 128)
    lin_case(256, 256, 256, bn=256)
    lin_case(1000, 320, 320)
    lin_case(1000, 320, 320, bias=True, res=True, alpha=0.5)
    lin_case(77, 1024, 640, bias=True)
    lin_case(4096, 1280, 1280, bias=True, res=True)
    lin_case(300, 320, 4, bias=True)          # tiny N, scalar store path
    lin_case(513, 192, 960, lda_pad=64)        # row-strided A
    lin_case(20000, 640, 1920)                 # many tiles (persistent loop, accumulator ping-pong)

    def geglu_case(m, k, inner, bn):
        def f():
            a = rnd(m, k).half()
            w = rnd(2 * inner, k, scale=k ** -0.5).half()
            b = rnd(2 * inner)
            wp, bp = ops.pack_geglu_weight(w, b, bn)
            out = ops.linear(a, wp, bias=bp, geglu=True, bn=bn)
            torch.cuda.synchronize()
            proj = (a.float() @ w.float().t() + b).half().float()
            v, gate = proj[:, :inner], proj[:, inner:]
            ref = v * torch.nn.functional.gelu(gate).half().float()
            report(f"[{impl}] geglu m{m} k{k} inner{inner} bn{bn}", rel_err(out, ref), 2e-3)
        run_case(f"geglu m{m}", f)

    geglu_case(256, 320, 1280, 256)
    geglu_case(1000, 640, 2560, 128)

    def conv_case(nimg, h, w_, c, n, bias=True, gb=False, res=False):
        def f():
            x = rnd(nimg, h, w_, c).half()
            wt = rnd(n, c, 3, 3, scale=(9 * c) ** -0.5).half()
            wp = wt.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous()
            b = rnd(n) if bias else None
            gbt = rnd(nimg, n).half() if gb else None
            r = rnd(nimg, h, w_, n).half() if res else None
            out = ops.conv2d_3x3(x, wp, bias=b, group_bias=gbt, residual=r)
            torch.cuda.synchronize()
            ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), wt.float(), b, padding=1)
            ref = ref.permute(0, 2, 3, 1)
            if gbt is not None:
                ref = ref.half().float() + gbt.float()[:, None, None, :]
            if r is not None:
                ref = ref.half().float() + r.float()
            report(f"[{impl}] conv3x3 n{nimg} {h}x{w_} c{c}->{n} gb{int(gb)} res{int(res)}", rel_err(out, ref), 2e-3)
        run_case(f"conv n{nimg} {h}x{w_}", f)

    conv_case(1, 8, 16, 64, 64)
    conv_case(2, 8, 16, 64, 64, gb=True, res=True)
    conv_case(2, 32, 32, 128, 320)
    conv_case(2, 22, 40, 128, 320, gb=True)
    conv_case(3, 11, 20, 192, 128)
    conv_case(2, 44, 80, 320, 640, gb=True, res=True)
    conv_case(1, 88, 160, 320, 4)

    def tconv_case(f_, hw, c, n, res=False):
        def f():
            x = rnd(f_, hw, c).half()
            wt = rnd(n, c, 3, scale=(3 * c) ** -0.5).half()
            wp = wt.permute(0, 2, 1).reshape(n, 3 * c).contiguous()
            b = rnd(n)
            r = rnd(f_, hw, n).half() if res else None
            out = ops.tconv3(x, wp, bias=b, residual=r)
            torch.cuda.synchronize()
            xr = x.float().permute(2, 0, 1).unsqueeze(0)  # [1, c, f, hw]
            ref = torch.nn.functional.conv2d(xr, wt.float().unsqueeze(-1), b, padding=(1, 0))
            ref = ref[0].permute(1, 2, 0)
            if r is not None:
                ref = ref.half().float() + r.float()
            report(f"[{impl}] tconv3 f{f_} hw{hw} c{c}->{n} res{int(res)}", rel_err(out, ref), 2e-3)
        run_case(f"tconv f{f_} hw{hw}", f)

    tconv_case(4, 256, 64, 64)
    tconv_case(8, 220, 128, 128, res=True)
    tconv_case(16, 32, 64, 96)
    tconv_case(16, 1024, 320, 320, res=True)
    tconv_case(1, 300, 64, 64)


def group_tapgemm_simt():
    group_tapgemm("simt")


GROUPS = {
    "tapgemm": group_tapgemm,
    "tapgemm_simt": group_tapgemm_simt,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--out", default="gpurun_out/gpu_check.json")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    if args.all:
        summary = {}
        for name in GROUPS:
            out = args.out.replace(".json", f".{name}.json")
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", name, "--out", out],
                                   timeout=args.timeout)
                summary[name] = {"rc": r.returncode, "s": round(time.time() - t0, 1)}
            except subprocess.TimeoutExpired:
                summary[name] = {"rc": "timeout", "s": round(time.time() - t0, 1)}
            print(f"== group {name}: {summary[name]}", flush=True)
        with open(args.out, "w") as fh:
            json.dump(summary, fh, indent=1)
        return 0
    print(f"device: {torch.cuda.get_device_name(0)}; torch {torch.__version__}", flush=True)
    GROUPS[args.group]()
    torch.cuda.synchronize()
    with open(args.out, "w") as fh:
        json.dump(RESULTS, fh, indent=1)
    nfail = sum(1 for r in RESULTS if not r["ok"])
    print(f"== {args.group}: {len(RESULTS) - nfail} pass, {nfail} fail", flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
