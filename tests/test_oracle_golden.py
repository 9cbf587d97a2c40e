"""The oracle (oracle/vgen_oracle.py) against the golden vectors frozen from the REAL reference
(oracle/make_golden.py).  Runs anywhere (no reference, no GPU): weights and inputs are regenerated
from their names."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import synth, vgen_oracle as vo
from oracle.cases import CASES, make_inputs


def _load(golden_dir, name):
    case = CASES[name]
    spec = [(k, tuple(s)) for k, s in json.load(open(os.path.join(golden_dir, f"{name}.spec.json")))]
    return case, synth.state_dict(spec, seed=case["seed"]), np.load(os.path.join(golden_dir, f"{name}.npz"))


def _maxrel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_schedules_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "schedules.npz"))
    b = vo.make_betas("cosine", 1000, True, cosine_s=0.008)
    assert np.array_equal(b.numpy(), g["cosine_zsnr.betas"])
    assert np.array_equal(vo.ddim_tables(b)["alphas_cumprod"].numpy(), g["cosine_zsnr.alphas_cumprod"])
    b2 = vo.make_betas("linear_sd", 1000, True, init_beta=0.00085, last_beta=0.012)
    assert np.array_equal(b2.numpy(), g["linear_sd_zsnr.betas"])
    assert np.array_equal(vo.ddim_tables(b2)["alphas_cumprod"].numpy(), g["linear_sd_zsnr.alphas_cumprod"])
    assert float(vo.ddim_tables(b)["alphas_cumprod"][999]) == 0.0           # zero terminal SNR
    for S in (50, 4, 20):
        assert np.array_equal(vo.ddim_steps(1000, S).numpy(), g[f"steps_{S}"])
    assert vo.ddim_steps(1000, 50)[:2].tolist() == [981, 961] and vo.ddim_steps(1000, 4).tolist() == [751, 501, 251, 1]


@pytest.mark.parametrize("name", ["t2v_tiny", "t2v_tiny_b2", "i2vgen_tiny", "vae_tiny"])
def test_oracle_matches_reference_golden(golden_dir, name):
    torch.set_grad_enabled(False)
    case, sd, gold = _load(golden_dir, name)
    inp = make_inputs(case)
    if case["kind"] == "t2v":
        out = vo.unet_t2v_forward(sd, inp["x"], inp["t"], inp["y"])
    elif case["kind"] == "i2vgen":
        out = vo.unet_i2vgen_forward(sd, inp["x"], inp["t"], inp["y"], inp["image"], inp["local_image"], inp["fps"])
    else:
        out = vo.vae_decode(sd, inp["z"])
    ref = torch.from_numpy(gold["out"])
    assert out.shape == ref.shape
    assert float(ref.std()) > 0.1, "golden output must not be degenerate (zero-init trap)"
    assert _maxrel(out, ref) < 5e-5


def test_oracle_ddim_loop_matches_reference_golden(golden_dir):
    torch.set_grad_enabled(False)
    case, sd, gold = _load(golden_dir, "t2v_tiny")
    inp = make_inputs(case)
    betas = vo.make_betas("cosine", 1000, True, cosine_s=0.008)
    kw = [{"y": inp["y"]}, {"y": inp["y_neg"]}]
    torch.manual_seed(123)
    lat = vo.ddim_sample_loop(inp["x"].clone(), lambda xt, t, **k: vo.unet_t2v_forward(sd, xt, t, **k), kw, betas,
                              case["ddim"]["guide_scale"], case["ddim"]["steps"])
    assert _maxrel(lat, torch.from_numpy(gold["ddim_latent"])) < 2e-4


def test_oracle_vae_encode_matches_reference_golden(golden_dir):
    torch.set_grad_enabled(False)
    case, sd, gold = _load(golden_dir, "vae_tiny")
    inp = make_inputs(case)
    mom = vo.vae_encode_moments(sd, inp["img"])
    assert _maxrel(mom, torch.from_numpy(gold["encode_moments"])) < 5e-5
    torch.manual_seed(case["encode"]["torch_seed"])       # the reference draws the posterior noise on the CPU generator
    z = vo.vae_encode_first_stage(sd, inp["img"], 0.18215)
    assert _maxrel(z, torch.from_numpy(gold["encode_z"])) < 5e-5
