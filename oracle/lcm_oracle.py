"""TEST INFRASTRUCTURE ONLY -- oracle for the LCM multistep consistency sampler of BASELINE config 3.

The reference engine (tools/inferences/inference_videolcm_entrance.py:171-179,233-255) drives
`diffusers.schedulers.LCMScheduler` (diffusers==0.26.3, tft2v_environment.yaml:60), a third-party dependency
that is NOT vendored under /root/reference and NOT installed here (no network).  PARITY PARTIALLY PINNED: this
file restates the published algorithm (arXiv:2310.04378, multistep consistency sampling; the scheduler's
documented behaviour) in plain torch.  Pinned against the reference's in-tree restatements of the scheduler
(tools/train/train_videolcm_t2v_entrance.py:129-176 -> oracle/make_golden_lcm.py -> tests/golden/lcm_pins.npz):
boundary_scalings, the x0 expression in sample_loop, the 50-step origin grid; and alphas_cumprod(True) bit-exact
against the reference's own rescale_zero_terminal_snr (tools/modules/diffusions/schedules.py:143-165).  PARITY
UNPINNED for the rest (inference timestep selection, re-noising update): no reference run or golden vector
exists.  What the tests additionally pin: the product scheduler (vgen_b200/lcm.py) equals this restatement, the
timestep list for the engine's arguments is [999, 759, 499, 259], zero terminal SNR holds (alphas_cumprod[999] == 0), and the
boundary condition c_skip -> 1, c_out -> 0 at t -> 0.

    sample_loop(noise, model, model_kwargs, steps) reproduces the engine loop with CFG off (its default).
"""
from __future__ import annotations

import numpy as np
import torch


def scaled_linear_betas(T=1000, beta_start=0.00085, beta_end=0.012):
    return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2


def rescale_zero_terminal_snr(betas):
    abar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
    abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    return 1 - torch.cat([abar[0:1], abar[1:] / abar[:-1]])


def alphas_cumprod(zero_snr=True):
    betas = scaled_linear_betas()
    if zero_snr:
        betas = rescale_zero_terminal_snr(betas)
    return torch.cumprod(1.0 - betas, dim=0)


def lcm_timesteps(num_inference_steps, original_steps=50, T=1000):
    k = T // original_steps
    origin = (np.arange(1, original_steps + 1) * k - 1)[::-1].copy()
    idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
    return [int(v) for v in origin[idx]]


def boundary_scalings(t, timestep_scaling=10.0, sigma_data=0.5):
    st = t * timestep_scaling
    return sigma_data ** 2 / (st ** 2 + sigma_data ** 2), st / (st ** 2 + sigma_data ** 2) ** 0.5


def sample_loop(noise, model, model_kwargs, steps=4):
    """v-prediction LCM sampling, CFG off: x0 = sqrt(abar) x - sqrt(1-abar) v; denoised = c_out x0 + c_skip x;
    x <- sqrt(abar_prev) denoised + sqrt(1-abar_prev) randn  (no re-noising after the last step)."""
    abar = alphas_cumprod(True)
    ts = lcm_timesteps(steps)
    x = noise
    for i, t in enumerate(ts):
        tt = torch.full((x.size(0),), t, device=x.device, dtype=x.dtype)
        v = model(x, tt, **model_kwargs).float()
        a_t = abar[t].to(x.device)
        x0 = a_t.sqrt() * x - (1 - a_t).sqrt() * v
        c_skip, c_out = boundary_scalings(float(t))
        den = c_out * x0 + c_skip * x
        if i != len(ts) - 1:
            a_p = abar[ts[i + 1]].to(x.device)
            x = a_p.sqrt() * den + (1 - a_p).sqrt() * torch.randn(x.shape, device=x.device, dtype=x.dtype)
        else:
            x = den
    return x
