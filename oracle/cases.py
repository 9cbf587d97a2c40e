"""TEST INFRASTRUCTURE ONLY -- the parity cases: constructor kwargs, weight seed and input recipe.

Small cases run the REAL reference on CPU in seconds (oracle/make_golden.py) and are replayed on the
GPU against the stored outputs.  Constructor kwargs are the reference's own (UNetSD_T2VBase /
UNetSD_I2VGen / AutoencoderKL); note the decoder's SpatialTransformers hard-code context_dim=1024
(tools/modules/unet/unet_t2v.py:180), so every case uses 1024-wide context tokens.
"""
from __future__ import annotations

import torch

from . import synth

_T2V_TINY = dict(in_dim=4, dim=64, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2], num_heads=2,
                 head_dim=64, num_res_blocks=1, attn_scales=[1.0, 0.5], dropout=0.1, temporal_attention=True,
                 temporal_attn_times=1, use_checkpoint=False, use_fps_condition=False, use_sim_mask=False)
_I2V_TINY = dict(_T2V_TINY, concat_dim=4)
_VAE_TINY = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32,
                               ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[], dropout=0.0), embed_dim=4)

_LCM_TINY = dict(_T2V_TINY, concat_dim=8, num_tokens=4)       # + config.video_compositions == ['text']
_HIGEN_TINY = dict(_T2V_TINY, context_embedding_depth=1, num_tokens=3)

# the effective constructor kwargs of the BASELINE configs
_FULL_UNET = dict(in_dim=4, dim=320, y_dim=1024, context_dim=1024, out_dim=4, dim_mult=[1, 2, 4, 4], num_heads=8,
                  head_dim=64, num_res_blocks=2, attn_scales=[1.0, 0.5, 0.25], dropout=0.1, temporal_attention=True,
                  temporal_attn_times=1, use_checkpoint=True, use_fps_condition=False, use_sim_mask=False)
_FULL_VAE = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                               ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)

FULL_CTORS = {
    "full_t2v": ("t2v", _FULL_UNET),
    "full_i2vgen": ("i2vgen", dict(_FULL_UNET, concat_dim=4)),
    "full_vae": ("vae", _FULL_VAE),
    # configs/videolcm_t2v_infer.yaml:46-66, sr600_infer.yaml:21-40, higen_infer.yaml (UNet section)
    "full_videolcm": ("videolcm", dict(_FULL_UNET, concat_dim=8, num_tokens=4)),
    "full_sr600": ("sr600", dict(_FULL_UNET)),
    "full_higen": ("higen", dict(_FULL_UNET, context_embedding_depth=2, num_tokens=16)),
}

LCM_CONFIG = dict(video_compositions=["text"], resolution=[448, 256])  # the `config` object UNetSD_VideoLCM reads (unet_videolcm.py:262-270)

CASES = {
    # name: kind, ctor kwargs, weight seed, input shape recipe
    "t2v_tiny": dict(kind="t2v", ctor=_T2V_TINY, seed=11, b=1, f=4, h=8, w=12, ntok=5, t=[481],
                     ddim=dict(steps=4, guide_scale=9.0)),
    "t2v_tiny_b2": dict(kind="t2v", ctor=_T2V_TINY, seed=12, b=2, f=3, h=10, w=6, ntok=7, t=[751, 21]),
    "i2vgen_tiny": dict(kind="i2vgen", ctor=_I2V_TINY, seed=13, b=1, f=4, h=8, w=12, ntok=5, t=[961],
                        ddim=dict(steps=4, guide_scale=9.0)),
    "videolcm_tiny": dict(kind="videolcm", ctor=_LCM_TINY, seed=15, b=1, f=4, h=8, w=12, ntok=5, t=[259]),
    "sr600_tiny": dict(kind="sr600", ctor=_T2V_TINY, seed=16, b=1, f=3, h=10, w=12, ntok=6, t=[600]),
    "higen_tiny": dict(kind="higen", ctor=_HIGEN_TINY, seed=17, b=2, f=3, h=8, w=6, ntok=5, t=[801, 333]),
    "higen_tiny_f1": dict(kind="higen", ctor=_HIGEN_TINY, seed=17, b=2, f=1, h=8, w=6, ntok=5, t=[401, 99]),
    "vae_tiny": dict(kind="vae", ctor=_VAE_TINY, seed=14, n=2, h=8, w=12, encode=dict(n=2, H=32, W=48, torch_seed=77)),
}


# Full-size parity cases: the BASELINE configs' real architectures and latent shapes.
# No golden file: truth = the (reference-pinned) oracle run in fp32 on the same GPU, tests/test_gpu_fullsize.py.
FULL_CASES = {
    "full_t2v": dict(kind="t2v", ctor=FULL_CTORS["full_t2v"][1], seed=21, b=1, f=8, h=32, w=32, ntok=77, t=[751]),
    "full_i2vgen": dict(kind="i2vgen", ctor=FULL_CTORS["full_i2vgen"][1], seed=22, b=1, f=16, h=88, w=160, ntok=77, t=[981]),
    "full_videolcm": dict(kind="videolcm", ctor=FULL_CTORS["full_videolcm"][1], seed=23, b=1, f=16, h=32, w=56, ntok=77, t=[759]),
    "full_sr600": dict(kind="sr600", ctor=FULL_CTORS["full_sr600"][1], seed=24, b=1, f=4, h=90, w=160, ntok=77, t=[600]),
    "full_higen_f1": dict(kind="higen", ctor=FULL_CTORS["full_higen"][1], seed=25, b=1, f=1, h=32, w=56, ntok=77, t=[981],
                          spec="full_higen"),
    "full_higen_f32": dict(kind="higen", ctor=FULL_CTORS["full_higen"][1], seed=25, b=1, f=32, h=32, w=56, ntok=77, t=[981],
                           spec="full_higen"),
    "full_vae": dict(kind="vae", ctor=FULL_CTORS["full_vae"][1], seed=26, n=2, h=88, w=160, z_scale=1.0 / 0.18215),
}


def make_inputs(case):
    """Deterministic inputs (numpy PCG64 keyed by tensor name, like the weights)."""
    s = case["seed"] + 1000
    if case["kind"] == "vae":
        d = {"z": synth.tensor("z", (case["n"], 4, case["h"], case["w"]), case.get("z_scale", 1.0), s)}
        if case.get("encode"):
            e = case["encode"]
            d["img"] = synth.tensor("img", (e["n"], 3, e["H"], e["W"]), 0.5, s).clamp(-1, 1)
        return d
    b, f, h, w, L = case["b"], case["f"], case["h"], case["w"], case["ntok"]
    d = {
        "x": synth.tensor("x", (b, 4, f, h, w), 1.0, s),
        "t": torch.tensor(case["t"], dtype=torch.long),
        "y": synth.tensor("y", (b, L, 1024), 1.0, s),
        "y_neg": synth.tensor("y_neg", (b, L, 1024), 1.0, s),
    }
    if case["kind"] == "i2vgen":
        d["image"] = synth.tensor("image", (b, 1, 1024), 1.0, s)
        li = synth.tensor("local_image", (b, 4, 1, h, w), 0.18215, s)
        d["local_image"] = li.repeat(1, 1, f, 1, 1)
        d["fps"] = torch.tensor([16] * b, dtype=torch.long)
    if case["kind"] == "higen":
        d["spat_prior"] = synth.tensor("spat_prior", (b, 4, h, w), 0.5, s)
        if f > 1:   # stage 2 (inference_higen_entrance.py:221-229): one motion factor per frame transition
            d["motion_cond"] = torch.tensor([[300 + 50 * i + 7 * j for j in range(f - 1)] for i in range(b)], dtype=torch.long)
        else:       # stage 1 (:198-203)
            d["motion_cond"] = torch.tensor([0] * b, dtype=torch.long)
        d["appearance_cond"] = synth.tensor("appearance_cond", (b, f, 32), 0.5, s).abs().clamp(0, 1)
    return d

# SR600 sampler pair (configs/sr600_infer.yaml:41-62, inference_sr600_entrance.py:253-280), shortened ladders
GAUSS_CASE = dict(
    unet_case="sr600_tiny",
    schedules=dict(
        reverse=dict(schedule="cosine", mean_type="v", schedule_param=dict(num_timesteps=1000, zero_terminal_snr=True)),
        forward=dict(schedule="logsnr_cosine_interp", mean_type="v",
                     schedule_param=dict(num_timesteps=1000, zero_terminal_snr=True, scale_min=2.0, scale_max=4.0))),
    noise_levels=600, reverse_steps=3, steps=4, guide_scale=9.0, guide_rescale=0.3, torch_seed=5,
)
