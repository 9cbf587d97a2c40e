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
from _helpers import oracle_call


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


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    torch.set_grad_enabled(False)
    case, sd, gold = _load(golden_dir, name)
    inp = make_inputs(case)
    out = oracle_call(case, sd, inp)
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


def test_fourier_filter_is_a_four_bin_update():
    """The SR600 skip filter only touches the 2x2 centre block of the shifted spectrum: the restated FFT
    filter equals x - (1-s) * Re(inverse DFT of the bins (ky, kx) in {0,-1}^2) -- the identity the CUDA
    kernel relies on (vgen_b200/csrc/variants.cu) -- for even and odd plane sizes."""
    torch.manual_seed(3)
    for hh, ww in ((5, 6), (12, 20), (7, 9), (2, 2)):
        x = torch.randn(2, 3, hh, ww, dtype=torch.float64)
        ref = vo.fourier_filter(x.float(), 1, 0.4).double()
        yy = torch.arange(hh, dtype=torch.float64).view(hh, 1) / hh
        xx = torch.arange(ww, dtype=torch.float64).view(1, ww) / ww
        low = torch.zeros_like(x)
        for ky in (0, -1):
            for kx in (0, -1):
                ph = 2 * np.pi * (ky * yy + kx * xx)
                basis = torch.complex(torch.cos(ph), torch.sin(ph))
                coef = (x * basis.conj()).sum(dim=(-2, -1), keepdim=True)
                low = low + (coef * basis).real
        mine = x - (1 - 0.4) * low / (hh * ww)
        assert float((mine - ref).abs().max()) < 1e-5


def test_gauss_oracle_matches_reference_golden(golden_dir):
    """SR600 sampler pair: sigma tables bit-exact, DDIM inversion and DPM-Solver++(2M) SDE latents vs the
    reference run (same seeded Brownian stub on both sides, oracle/brownian.py)."""
    from oracle import gauss_oracle as go
    from oracle.cases import GAUSS_CASE as gc
    torch.set_grad_enabled(False)
    g = np.load(os.path.join(golden_dir, "gauss.npz"))
    sig = {}
    for name, kw in gc["schedules"].items():
        sig[name] = go.sigma_schedule(kw["schedule"], **kw["schedule_param"])
        assert np.array_equal(sig[name].numpy(), g[f"sigmas.{name}"])
    case, sd, _ = _load(golden_dir, gc["unet_case"])
    inp = make_inputs(case)
    fn = lambda xt, t, **k: vo.unet_sr600_forward(sd, xt, t, **k)  # noqa: E731
    rev = go.GaussOracle(sig["reverse"], "v").ddim_reverse_sample_loop(
        inp["x"], fn, {"y": inp["y_neg"]}, ddim_timesteps=gc["reverse_steps"], reverse_steps=gc["noise_levels"])
    assert _maxrel(rev, torch.from_numpy(g["reverse_latent"])) < 5e-5
    fwd = go.GaussOracle(sig["forward"], "v")
    lad = fwd.sample_sigmas(gc["steps"], gc["noise_levels"] - 1, 0, "trailing", True)
    assert np.array_equal(lad[:-1].numpy(), g["ladder"][:gc["steps"]]) and float(lad[-1]) == 0.0
    torch.manual_seed(gc["torch_seed"])
    lat = fwd.sample_dpmpp_2m_sde(torch.from_numpy(g["reverse_latent"]), fn, [{"y": inp["y"]}, {"y": inp["y_neg"]}],
                                  gc["guide_scale"], gc["guide_rescale"], steps=gc["steps"], t_max=gc["noise_levels"] - 1,
                                  t_min=0, discretization="trailing")
    assert _maxrel(lat, torch.from_numpy(g["sample_latent"])) < 2e-4


def test_video_oracle_matches_reference_frames(golden_dir):
    """oracle/video_oracle.py vs the frames the REAL save_i2vgen_video_safe handed to its encoder
    (tests/golden/video_out.npz, oracle/make_golden_video.py): bit-exact, including the dropped grey last frame."""
    import numpy as np
    from oracle import make_golden_video as mg, video_oracle as vo_
    g = np.load(os.path.join(golden_dir, "video_out.npz"))
    for name in mg.CASES:
        v, mean, std = mg.make_video(name)
        frames = vo_.drop_anomalous_last_frame(vo_.frames_uint8(v.numpy(), mean, std))
        gold = g[name + "_frames"]
        assert len(frames) == gold.shape[0], name
        assert np.array_equal(np.stack(frames), gold), name
    assert g["grey_last_frames"].shape[0] == int(g["grey_last_nframes_in"]) - 1


def test_clip_oracle_matches_reference_golden(golden_dir):
    """oracle/clip_oracle.py vs outputs frozen from the reference's vendored open_clip CLIP driven like
    tools/modules/clip_embedder.py:183-212 (tests/golden/clip_tiny.npz, oracle/make_golden_clip.py)."""
    from oracle import clip_oracle as co
    from oracle.make_golden_clip import TINY
    g = np.load(os.path.join(golden_dir, "clip_tiny.npz"))
    spec = [(k, tuple(s)) for k, s in json.load(open(os.path.join(golden_dir, "clip_tiny.spec.json")))]
    sd = synth.state_dict([s for s in spec if len(s[1]) > 0], seed=31)
    for k in ("positional_embedding", "visual.positional_embedding", "visual.class_embedding", "token_embedding.weight"):
        sd[k] = synth.tensor(k, dict(spec)[k], 0.02, 31)
    tokens = torch.from_numpy(g["tokens"])
    image = synth.tensor("clip_image", (2, 3, 56, 56), 1.0, 31)
    for layer in ("last", "penultimate"):
        xt, x = co.encode_text(sd, tokens, TINY["text_cfg"]["heads"], layer)
        assert float((xt - torch.from_numpy(g[f"{layer}_xt"])).abs().max()) < 1e-4 * float(np.abs(g[f"{layer}_xt"]).max())
        assert float((x - torch.from_numpy(g[f"{layer}_x"])).abs().max()) < 1e-4 * float(np.abs(g[f"{layer}_x"]).max())
    xi = co.encode_image(sd, image, TINY["vision_cfg"]["head_width"])
    assert float((xi - torch.from_numpy(g["xi"])).abs().max()) < 1e-4 * float(np.abs(g["xi"]).max())


def test_lcm_oracle_matches_reference_in_tree_pieces(golden_dir):
    """Partial pin of the LCM sampler (diffusers absent): boundary scalings, the step-4 x0 formula and the 50-step DDIM
    grid, frozen from the reference's own in-tree restatements (tools/train/train_videolcm_t2v_entrance.py:129-176,
    oracle/make_golden_lcm.py) -- oracle and product host math against them."""
    from oracle import lcm_oracle as lo
    from vgen_b200.lcm import LCMScheduler
    g = np.load(os.path.join(golden_dir, "lcm_pins.npz"))
    ts = g["timesteps"].tolist()
    s = LCMScheduler(prediction_type="v_prediction", beta_schedule="scaled_linear", clip_sample=False,
                     timestep_spacing="linspace", rescale_betas_zero_snr=True)
    for i, t in enumerate(ts):
        for got in (lo.boundary_scalings(float(t)), s.boundary_scalings(t)):
            assert got[0] == pytest.approx(float(g["c_skip"][i]), rel=1e-12, abs=0)
            assert got[1] == pytest.approx(float(g["c_out"][i]), rel=1e-12, abs=0)
    # x0 (LCMScheduler.step, step 4): the oracle's in-loop expression, v-prediction and epsilon forms
    abar = lo.alphas_cumprod(True).double()
    shape = g["x0_v"].shape
    x = synth.tensor("lcm_sample", shape, 1.0, 3).double()
    v = synth.tensor("lcm_model_output", shape, 1.0, 4).double()
    for i, t in enumerate(ts):
        a = abar[t]
        x0 = a.sqrt() * x[i] - (1 - a).sqrt() * v[i]                     # lcm_oracle.sample_loop, v-prediction
        assert np.array_equal(x0.numpy(), g["x0_v"][i])
        if i:
            x0e = (x[i] - (1 - a).sqrt() * v[i]) / a.sqrt()
            assert np.allclose(x0e.numpy(), g["x0_eps"][i - 1], rtol=1e-12, atol=0)
    # the grid the consistency model was distilled on == the origin grid both schedulers subsample
    assert g["ddim_grid"].tolist() == sorted(lo.lcm_timesteps(50)) == [20 * k + 19 for k in range(50)]
    s.set_timesteps(50)
    assert sorted(s.timesteps.tolist()) == g["ddim_grid"].tolist()
    assert np.array_equal(g["ddim_abar"], abar.numpy()[g["ddim_grid"]])
    assert set(lo.lcm_timesteps(4)) <= set(g["ddim_grid"].tolist())
    # zero-terminal-SNR rescale: the reference's own implementation (schedules.py:143-165) on the same fp32 betas
    assert np.array_equal(lo.alphas_cumprod(True).numpy(), g["abar_zero_snr"])
    assert np.array_equal(s.alphas_cumprod.numpy(), g["abar_zero_snr"]) and g["abar_zero_snr"][-1] == 0.0
