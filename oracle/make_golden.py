"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the REAL reference into tests/golden/.

Run in the build container (the reference is mounted read-only at /root/reference):

    python -m oracle.make_golden

For every case it (1) builds the reference class on CPU, (2) loads a synthetic state_dict made by
oracle/synth.py from the reference's own (name, shape) list, (3) runs the reference forward in fp32,
(4) checks the oracle restatement (oracle/vgen_oracle.py) against it, and (5) stores the parameter spec
and the reference output.  The GPU box has no reference: tests regenerate weights/inputs from the
recipes here and compare with the stored outputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refload, synth, vgen_oracle as vo  # noqa: E402
from oracle.cases import CASES, LCM_CONFIG, make_inputs  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _maxrel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def build_variant(ref, kind, ctor):
    """The a21 model variants: constructor calls as the inference entrances make them."""
    if kind == "videolcm":
        from easydict import EasyDict
        return ref.UNetSD_VideoLCM(config=EasyDict(**LCM_CONFIG), **ctor)
    if kind == "sr600":
        return ref.UNetSD_SR600(**ctor)
    if kind == "higen":
        return ref.UNetSD_HiGen(**ctor)
    raise ValueError(kind)


def main():
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    ref = refload.load()
    os.makedirs(GOLD, exist_ok=True)
    report = {}

    # ---- schedules / DDIM tables: bit-exact fp64
    sched = {}
    for name, kw in (("cosine_zsnr", dict(schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True))),
                     ("linear_sd_zsnr", dict(schedule="linear_sd", schedule_param=dict(num_timesteps=1000, init_beta=0.00085, last_beta=0.012, zero_terminal_snr=True)))):
        d = ref.DiffusionDDIM(mean_type="v", var_type="fixed_small", **kw)
        sp = dict(kw["schedule_param"])
        zs = sp.pop("zero_terminal_snr")
        mine = vo.make_betas(kw["schedule"], zero_terminal_snr_flag=zs, **sp)
        assert torch.equal(mine, d.betas), name
        tab = vo.ddim_tables(mine)
        for k in tab:
            assert torch.equal(tab[k], getattr(d, k)), (name, k)
        sched[name + ".betas"] = d.betas.numpy()
        sched[name + ".alphas_cumprod"] = d.alphas_cumprod.numpy()
    for S in (50, 4, 20):
        steps = (1 + torch.arange(0, 1000, 1000 // S)).clamp(0, 999).flip(0)
        assert torch.equal(steps, vo.ddim_steps(1000, S))
        sched[f"steps_{S}"] = steps.numpy()
    np.savez_compressed(os.path.join(GOLD, "schedules.npz"), **sched)
    report["schedules"] = "bit-exact"

    # ---- model cases
    for cname, case in CASES.items():
        kind = case["kind"]
        if kind == "t2v":
            m = ref.UNetSD_T2VBase(**case["ctor"]).eval()
        elif kind == "i2vgen":
            m = ref.UNetSD_I2VGen(**case["ctor"]).eval()
        elif kind == "vae":
            m = ref.AutoencoderKL(**case["ctor"]).eval()
        else:
            m = build_variant(ref, kind, case["ctor"]).eval()
        spec = synth.spec_of(m)
        sd = synth.state_dict(spec, seed=case["seed"])
        m.load_state_dict(sd, strict=True)
        inp = make_inputs(case)
        if kind == "t2v":
            out = m(inp["x"], inp["t"], y=inp["y"])
            mine = vo.unet_t2v_forward(sd, inp["x"], inp["t"], inp["y"])
        elif kind == "i2vgen":
            out = m(inp["x"], inp["t"], y=inp["y"], image=inp["image"], local_image=inp["local_image"], fps=inp["fps"])
            mine = vo.unet_i2vgen_forward(sd, inp["x"], inp["t"], inp["y"], inp["image"], inp["local_image"], inp["fps"])
        elif kind == "videolcm":
            out = m(inp["x"], inp["t"], y=inp["y"])
            mine = vo.unet_videolcm_forward(sd, inp["x"], inp["t"], inp["y"])
        elif kind == "sr600":
            out = m(inp["x"], inp["t"], inp["y"])
            mine = vo.unet_sr600_forward(sd, inp["x"], inp["t"], inp["y"])
        elif kind == "higen":
            hk = dict(spat_prior=inp["spat_prior"], motion_cond=inp["motion_cond"], appearance_cond=inp["appearance_cond"])
            out = m(inp["x"], inp["t"], y=inp["y"], **hk)
            mine = vo.unet_higen_forward(sd, inp["x"], inp["t"], inp["y"], **hk)
        else:
            out = m.decode(inp["z"])
            mine = vo.vae_decode(sd, inp["z"])
        err = _maxrel(mine, out)
        assert err < 2e-5, (cname, err)
        arrays = {"out": out.numpy()}
        extra = {}
        if kind == "vae" and case.get("encode"):
            torch.manual_seed(case["encode"]["torch_seed"])
            zr = m.encode_firsr_stage(inp["img"], 0.18215)
            torch.manual_seed(case["encode"]["torch_seed"])
            zo = vo.vae_encode_first_stage(sd, inp["img"], 0.18215)
            e3 = _maxrel(zo, zr)
            assert e3 < 2e-5, (cname, "encode", e3)
            arrays["encode_z"] = zr.numpy()
            arrays["encode_moments"] = m.quant_conv(m.encoder(inp["img"])).numpy()
            extra["encode_err"] = e3
        if case.get("ddim"):
            # full sampler on the reference: DiffusionDDIM.ddim_sample_loop with CFG
            dd = case["ddim"]
            diff = ref.DiffusionDDIM(schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True),
                                     mean_type="v", var_type="fixed_small")
            kw = [{"y": inp["y"]}, {"y": inp["y_neg"]}] if kind == "t2v" else \
                [{"y": inp["y"], "image": inp["image"], "local_image": inp["local_image"], "fps": inp["fps"]},
                 {"y": inp["y_neg"], "image": torch.zeros_like(inp["image"]), "local_image": inp["local_image"], "fps": inp["fps"]}]
            torch.manual_seed(123)
            lat = diff.ddim_sample_loop(inp["x"].clone(), m, kw, guide_scale=dd["guide_scale"], ddim_timesteps=dd["steps"], eta=0.0)
            betas = vo.make_betas("cosine", 1000, True, cosine_s=0.008)
            if kind == "t2v":
                fn = lambda xt, t, **k: vo.unet_t2v_forward(sd, xt, t, **k)  # noqa: E731
            else:
                fn = lambda xt, t, **k: vo.unet_i2vgen_forward(sd, xt, t, **k)  # noqa: E731
            torch.manual_seed(123)
            mine_lat = vo.ddim_sample_loop(inp["x"].clone(), fn, kw, betas, dd["guide_scale"], dd["steps"])
            e2 = _maxrel(mine_lat, lat)
            assert e2 < 1e-4, (cname, "ddim", e2)
            arrays["ddim_latent"] = lat.numpy()
            extra["ddim_err"] = e2
        np.savez_compressed(os.path.join(GOLD, f"{cname}.npz"), **arrays)
        with open(os.path.join(GOLD, f"{cname}.spec.json"), "w") as fh:
            json.dump([[k, list(s)] for k, s in spec], fh)
        report[cname] = {"oracle_vs_reference_maxrel": err, "params": int(sum(int(np.prod(s)) for _, s in spec)), **extra}
        print(cname, report[cname], flush=True)

    # ---- GaussianDiffusion / DiffusionDDIMSR (SR600 sampler pair) on the reference, with the torchsde stub
    from oracle import gauss_oracle as go
    from oracle.cases import GAUSS_CASE as gc
    gold = {}
    for name, kw in gc["schedules"].items():
        ref_sig = ref.schedules.sigma_schedule(kw["schedule"], **kw["schedule_param"])
        assert torch.equal(ref_sig, go.sigma_schedule(kw["schedule"], **kw["schedule_param"])), name
        gold[f"sigmas.{name}"] = ref_sig.numpy()
    case = CASES[gc["unet_case"]]
    m = ref.UNetSD_SR600(**case["ctor"]).eval()
    sd = synth.state_dict(synth.spec_of(m), seed=case["seed"])
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(case)
    from easydict import EasyDict
    dsr = ref.diffusion_ddim.DiffusionDDIMSR(EasyDict(gc["schedules"]["reverse"]), EasyDict(gc["schedules"]["forward"]))
    fn = lambda xt, t, **k: vo.unet_sr600_forward(sd, xt, t, **k)  # noqa: E731
    rev = dsr.reverse_diffusion.ddim_reverse_sample_loop(x0=inp["x"], model=m, model_kwargs={"y": inp["y_neg"]},
                                                         ddim_timesteps=gc["reverse_steps"], reverse_steps=gc["noise_levels"])
    o_rev = go.GaussOracle(dsr.reverse_diffusion.sigmas, "v").ddim_reverse_sample_loop(
        inp["x"], fn, {"y": inp["y_neg"]}, ddim_timesteps=gc["reverse_steps"], reverse_steps=gc["noise_levels"])
    kw = [{"y": inp["y"]}, {"y": inp["y_neg"]}]
    skw = dict(guide_scale=gc["guide_scale"], guide_rescale=gc["guide_rescale"], steps=gc["steps"],
               t_max=gc["noise_levels"] - 1, t_min=0, discretization="trailing")
    torch.manual_seed(gc["torch_seed"])
    lat = dsr.forward_diffusion.sample(noise=rev, model=m, model_kwargs=kw, solver="dpmpp_2m_sde", **skw)
    torch.manual_seed(gc["torch_seed"])
    o_lat = go.GaussOracle(dsr.forward_diffusion.sigmas, "v").sample_dpmpp_2m_sde(rev, fn, kw, **skw)
    e_rev, e_lat = _maxrel(o_rev, rev), _maxrel(o_lat, lat)
    assert e_rev < 2e-5 and e_lat < 1e-4, (e_rev, e_lat)
    gold["ladder"] = dsr.forward_diffusion._t_to_sigma(torch.arange(gc["noise_levels"] - 1, -1, -(gc["noise_levels"] / (gc["steps"] + 1))).clamp_(0, gc["noise_levels"] - 1)).numpy()
    gold["reverse_latent"], gold["sample_latent"] = rev.numpy(), lat.numpy()
    np.savez_compressed(os.path.join(GOLD, "gauss.npz"), **gold)
    report["gauss"] = {"reverse_err": e_rev, "sample_err": e_lat}
    print("gauss", report["gauss"], flush=True)

    # ---- full-size parameter specs (names/shapes only) for strict state_dict compatibility tests
    from oracle.cases import FULL_CTORS
    for name, (kind, ctor) in FULL_CTORS.items():
        with torch.device("meta"):
            if kind in ("t2v", "i2vgen", "vae"):
                m = {"t2v": ref.UNetSD_T2VBase, "i2vgen": ref.UNetSD_I2VGen, "vae": ref.AutoencoderKL}[kind](**ctor)
            else:
                m = build_variant(ref, kind, ctor)
        spec = synth.spec_of(m)
        with open(os.path.join(GOLD, f"{name}.spec.json"), "w") as fh:
            json.dump([[k, list(s)] for k, s in spec], fh)
        report[name] = {"tensors": len(spec), "params": int(sum(int(np.prod(s)) for _, s in spec))}
        print(name, report[name], flush=True)

    with open(os.path.join(GOLD, "REPORT.json"), "w") as fh:
        json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
