"""Model-level parity report on a GPU: vgen_b200 (CUDA kernels) vs the oracle (fp32 and fp16-autocast
PyTorch restatement of the reference) vs the golden vectors frozen from the real reference.

    python tools/gpu_model_check.py            # all small cases
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth, vgen_oracle as vo  # noqa: E402
from oracle.cases import CASES, make_inputs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def errs(a, b):
    a, b = a.float(), b.float()
    l2 = ((a - b).norm() / (b.norm() + 1e-12)).item()
    mx = ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
    return l2, mx


def load_case(name):
    case = CASES[name]
    spec = [(k, tuple(s)) for k, s in json.load(open(os.path.join(GOLD, f"{name}.spec.json")))]
    sd = synth.state_dict(spec, seed=case["seed"])
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    return case, sd, gold


def build_model(case, sd):
    import vgen_b200
    if case["kind"] == "t2v":
        m = vgen_b200.UNetSD_T2VBase(**case["ctor"])
    elif case["kind"] == "i2vgen":
        m = vgen_b200.UNetSD_I2VGen(**case["ctor"])
    else:
        m = vgen_b200.AutoencoderKL(**case["ctor"])
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def run_case(name, report):
    case, sd, gold = load_case(name)
    inp = {k: v.cuda() for k, v in make_inputs(case).items()}
    sdg = {k: v.cuda() for k, v in sd.items()}
    m = build_model(case, sd)
    kind = case["kind"]
    with torch.no_grad():
        if kind == "t2v":
            mine = m(inp["x"], inp["t"], y=inp["y"])
            ofn = lambda: vo.unet_t2v_forward(sdg, inp["x"], inp["t"], inp["y"])  # noqa: E731
        elif kind == "i2vgen":
            mine = m(inp["x"], inp["t"], y=inp["y"], image=inp["image"], local_image=inp["local_image"], fps=inp["fps"])
            ofn = lambda: vo.unet_i2vgen_forward(sdg, inp["x"], inp["t"], inp["y"], inp["image"], inp["local_image"], inp["fps"])  # noqa: E731
        else:
            mine = m.decode(inp["z"])
            ofn = lambda: vo.vae_decode(sdg, inp["z"])  # noqa: E731
        torch.cuda.synchronize()
        o32 = ofn()
        with torch.autocast("cuda", dtype=torch.float16):
            o16 = ofn()
    g = torch.from_numpy(gold["out"]).cuda()
    r = {"mine_vs_golden": errs(mine, g), "mine_vs_oracle32": errs(mine, o32), "autocast_vs_oracle32": errs(o16, o32),
         "oracle32gpu_vs_golden": errs(o32, g), "mine_vs_autocast": errs(mine, o16),
         "finite": bool(torch.isfinite(mine.float()).all()), "out_std": float(g.std())}
    report[name] = r
    print(name, json.dumps(r), flush=True)

    if case.get("ddim") and kind in ("t2v", "i2vgen"):
        import vgen_b200
        dd = case["ddim"]
        diff = vgen_b200.DiffusionDDIM(schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True),
                                       mean_type="v", var_type="fixed_small")
        if kind == "t2v":
            kw = [{"y": inp["y"]}, {"y": inp["y_neg"]}]
        else:
            kw = [{"y": inp["y"], "image": inp["image"], "local_image": inp["local_image"], "fps": inp["fps"]},
                  {"y": inp["y_neg"], "image": torch.zeros_like(inp["image"]), "local_image": inp["local_image"], "fps": inp["fps"]}]
        torch.manual_seed(123)
        lat = diff.ddim_sample_loop(inp["x"].clone(), m, kw, guide_scale=dd["guide_scale"], ddim_timesteps=dd["steps"], eta=0.0)
        torch.cuda.synchronize()
        gl = torch.from_numpy(gold["ddim_latent"]).cuda()
        betas = vo.make_betas("cosine", 1000, True, cosine_s=0.008)
        fn = (lambda xt, t, **k: vo.unet_t2v_forward(sdg, xt, t, **k)) if kind == "t2v" else \
            (lambda xt, t, **k: vo.unet_i2vgen_forward(sdg, xt, t, **k))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            lat16 = vo.ddim_sample_loop(inp["x"].clone(), fn, kw, betas, dd["guide_scale"], dd["steps"], autocast_cfg=True)
        r2 = {"ddim_mine_vs_golden": errs(lat, gl), "ddim_autocast_vs_golden": errs(lat16, gl), "ddim_mine_vs_autocast": errs(lat, lat16)}
        report[name + ".ddim"] = r2
        print(name + ".ddim", json.dumps(r2), flush=True)


def main():
    report = {}
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t0 = time.time()
        try:
            run_case(name, report)
        except Exception as e:  # noqa: BLE001
            report[name] = {"exc": repr(e)}
            print(name, "EXC", repr(e), flush=True)
            traceback.print_exc()
        print(f"  ({name}: {time.time() - t0:.1f}s)", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "model_check.json"), "w") as fh:
        json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
