"""B200-native stand-in for `diffusers.schedulers.LCMScheduler` as the VideoLCM engine uses it
(tools/inferences/inference_videolcm_entrance.py:171-179,233-255; BASELINE config 3):

    diffusion = LCMScheduler(prediction_type="v_prediction", beta_schedule="scaled_linear", clip_sample=False,
                             timestep_spacing="linspace", rescale_betas_zero_snr=True)
    diffusion.set_timesteps(4, device=model.device)
    for t in diffusion.timesteps:
        latent_model_input = diffusion.scale_model_input(latents, t)
        noise_pred = model(latent_model_input, t.repeat(b).to(latents.dtype), t_w=None, **kw)
        latents = diffusion.step(noise_pred, t, latents, return_dict=False)[0]

diffusers is a third-party dependency of that engine (diffusers==0.26.3, tft2v_environment.yaml:60) and is
not installed here, so this file restates the published algorithm of its LCMScheduler (latent consistency
models, arXiv:2310.04378; multistep consistency sampling): **parity partially pinned** (boundary scalings, x0
formula and the 50-step grid against tools/train/train_videolcm_t2v_entrance.py:129-176, the zero-SNR rescale against
tools/modules/diffusions/schedules.py:143-165), timestep selection and re-noising unpinned -- see
DESIGN.md §4.  Points to
re-verify against a real diffusers 0.26.3: the timestep list for (50 original steps, 4 inference steps) and
the treatment of alphas_cumprod[999] == 0 after the zero-terminal-SNR rescale.

Host side: schedule tables (fp32 like diffusers), timestep selection, per-step scalars.  Device side: the
same step kernels as GaussianDiffusion (vgen_gauss_x0, vgen_lincomb_f32); the re-noising draw is
`torch.randn` on the sample's device (diffusers' `randn_tensor` with generator=None).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def _rescale_zero_terminal_snr(betas):
    """diffusers' rescale_zero_terminal_snr (arXiv:2305.08891, alg. 1), fp32."""
    alphas = 1.0 - betas
    abar_sqrt = torch.cumprod(alphas, dim=0).sqrt()
    a0, aT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
    abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return 1 - alphas


class LCMScheduler(object):
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 original_inference_steps=50, clip_sample=False, set_alpha_to_one=True, prediction_type="epsilon",
                 timestep_spacing="leading", timestep_scaling=10.0, rescale_betas_zero_snr=False, **kwargs):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"vgen_b200 LCMScheduler: beta_schedule '{beta_schedule}'")
        if clip_sample:
            raise NotImplementedError("vgen_b200 LCMScheduler: clip_sample is not used by the VideoLCM engine")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(prediction_type)
        if rescale_betas_zero_snr:
            betas = _rescale_zero_terminal_snr(betas)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        self.prediction_type = prediction_type
        self.timestep_scaling = timestep_scaling
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._step_index = None

    # ---- host: timestep selection --------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps, device=None, original_inference_steps=None, strength=1.0):
        original = original_inference_steps or self.original_inference_steps
        if original > self.num_train_timesteps or num_inference_steps > original:
            raise ValueError("LCMScheduler.set_timesteps: inconsistent step counts")
        k = self.num_train_timesteps // original
        origin = np.asarray(list(range(1, int(original * strength) + 1))) * k - 1      # 19, 39, ..., 999
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(origin[idx].astype(np.int64)).to(device=device)
        self._host_timesteps = [int(v) for v in origin[idx]]
        self._step_index = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def boundary_scalings(self, timestep):
        """c_skip, c_out of the consistency parameterisation (sigma_data = 0.5, timestep * timestep_scaling)."""
        sigma_data = 0.5
        st = float(timestep) * self.timestep_scaling
        return sigma_data ** 2 / (st ** 2 + sigma_data ** 2), st / math.sqrt(st ** 2 + sigma_data ** 2)

    # ---- one multistep-consistency step --------------------------------------------------------------
    @torch.no_grad()
    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("LCMScheduler.step: call set_timesteps first")
        if self._step_index is None:
            self._step_index = 0
        i = self._step_index
        t = self._host_timesteps[i]                     # host copy: no device->host read of `timestep`
        t_prev = self._host_timesteps[i + 1] if i + 1 < len(self._host_timesteps) else t
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t_prev] if t_prev >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.boundary_scalings(t)
        x = sample if (sample.dtype == torch.float32 and sample.is_contiguous()) else sample.float().contiguous()
        out = model_output if (model_output.dtype == torch.float16 and model_output.is_contiguous()) \
            else model_output.to(torch.float16).contiguous()
        pred = {"v_prediction": "v", "epsilon": "eps", "sample": "x0"}[self.prediction_type]
        x0 = ops.gauss_x0(x, out, float(a_t.sqrt()), float((1 - a_t).sqrt()), pred)
        denoised = ops.lincomb_f32([(c_out, x0), (c_skip, x)])
        if i != self.num_inference_steps - 1:
            noise = torch.randn(x.shape, generator=generator, device=x.device, dtype=x.dtype)
            prev = ops.lincomb_f32([(float(a_prev.sqrt()), denoised), (float((1 - a_prev).sqrt()), noise)])
        else:
            prev = denoised
        self._step_index = i + 1
        if return_dict:
            import types
            return types.SimpleNamespace(prev_sample=prev, denoised=denoised)
        return (prev, denoised)
