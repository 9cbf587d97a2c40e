"""Full-architecture smoke of the a21 model variants at BASELINE-config-like latent sizes (random non-zero weights):
checks that every tensor map / tile shape of the real 1.4 B-parameter layouts encodes and that outputs are finite.
Parity itself is established on the tiny cases against the reference golden vectors (tests/test_gpu_parity.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vgen_b200  # noqa: E402
from oracle.cases import FULL_CTORS, LCM_CONFIG  # noqa: E402


def build(kind, ctor):
    cls = {"videolcm": vgen_b200.UNetSD_VideoLCM, "sr600": vgen_b200.UNetSD_SR600, "higen": vgen_b200.UNetSD_HiGen}[kind]
    m = cls(config=dict(LCM_CONFIG), **ctor) if kind == "videolcm" else cls(**ctor)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            if float(p.abs().sum()) == 0.0:
                p.normal_(0, 0.02, generator=g)
    return m.cuda().eval()


def main():
    g = torch.Generator().manual_seed(2)
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
    runs = [("videolcm", (1, 4, 16, 32, 56)), ("sr600", (1, 4, 8, 90, 160)), ("higen", (1, 4, 1, 32, 56)), ("higen", (1, 4, 8, 32, 56))]
    models = {}
    for kind, shape in runs:
        if kind not in models:
            t0 = time.time()
            models[kind] = build(kind, FULL_CTORS["full_" + kind][1])
            print(f"built {kind} in {time.time() - t0:.1f}s", flush=True)
        m = models[kind]
        b, c, f, h, w = shape
        x, t, y = rnd(*shape), torch.tensor([500] * b, device="cuda"), rnd(b, 77, 1024)
        kw = {}
        if kind == "higen":
            kw = dict(spat_prior=rnd(b, 4, h, w), appearance_cond=torch.rand(b, f, 32, generator=g).cuda(),
                      motion_cond=(torch.full((b, f - 1), 500, dtype=torch.long) if f > 1 else torch.zeros(b, dtype=torch.long)).cuda())
        out = m(x, t, y, **kw) if kind == "sr600" else m(x, t, y=y, **kw)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = m(x, t, y, **kw) if kind == "sr600" else m(x, t, y=y, **kw)
        e.record()
        torch.cuda.synchronize()
        ok = bool(torch.isfinite(out.float()).all()) and tuple(out.shape) == (b, 4, f, h, w) and float(out.float().std()) > 0
        print(f"{'PASS' if ok else 'FAIL'} {kind} latent {shape}: forward {s.elapsed_time(e):.1f} ms, out std {float(out.float().std()):.3f}", flush=True)
        if kind != "higen" or f > 1:
            del models[kind]
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
