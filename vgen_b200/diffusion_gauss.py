"""B200-native GaussianDiffusion / DiffusionDDIMSR: the sampler pair of the SR600 pipeline
(tools/modules/diffusions/diffusion_gauss.py:145-499, diffusion_ddim.py:19-25)

    diffusion.reverse_diffusion.ddim_reverse_sample_loop(x0=..., ddim_timesteps=30, reverse_steps=T)
    diffusion.forward_diffusion.sample(noise=..., guide_scale=9.0, guide_rescale=0.3, solver='dpmpp_2m_sde',
                                       steps=30, t_max=T-1, t_min=0, discretization='trailing')
                                                                   (inference_sr600_entrance.py:256-280)

Host side (torch CPU tensors, the reference's own dtypes): sigma schedules, the sigma ladder, the
sigma <-> fractional-timestep interpolation and the per-step scalar coefficients.  Device side: three
kernels of libvgen_b200.so per step (vgen_cfg_combine, vgen_gauss_x0, vgen_lincomb_f32) around the UNet
forwards.  All batch entries share the timestep (true for both entry points).  Training-time pieces
(losses), the karras ladders and the solvers the reference itself comments out are absent.
"""
from __future__ import annotations

import math
import random

import torch

from . import brownian, ops
from .diffusion import _cosine_betas, _linear_sd_betas, _rescale_zero_terminal_snr, cfg_forward

BROWNIAN_TREE = None   # override hook: a class with the torchsde.BrownianTree interface; None = default_tree_cls()


# ---------------------------------------------------------------------------------------- schedules
def _logsnr_cosine(n, logsnr_min, logsnr_max):
    t_min = math.atan(math.exp(-0.5 * logsnr_min))
    t_max = math.atan(math.exp(-0.5 * logsnr_max))
    t = torch.linspace(1, 0, n)
    return -2 * torch.log(torch.tan(t_min + t * (t_max - t_min)))


def _logsnr_cosine_interp_sigmas(num_timesteps, scale_min=2, scale_max=4, logsnr_min=-15, logsnr_max=15, **kw):
    """schedules.py:52-60,114-140."""
    t = torch.linspace(1, 0, num_timesteps)
    lo = _logsnr_cosine(num_timesteps, logsnr_min, logsnr_max) + 2 * math.log(1 / scale_min)
    hi = _logsnr_cosine(num_timesteps, logsnr_min, logsnr_max) + 2 * math.log(1 / scale_max)
    return torch.sqrt(torch.sigmoid(-(t * lo + (1 - t) * hi)))


def sigma_schedule(schedule="cosine", num_timesteps=1000, zero_terminal_snr=False, **kwargs):
    """schedules.py:24-43, including its quirk: the beta-domain zero-terminal-SNR rescale is applied to the
    sigma table whenever max(sigma) is further than 1e-4 from 1 (it is not for the released configs)."""
    if schedule == "logsnr_cosine_interp":
        sigma = _logsnr_cosine_interp_sigmas(num_timesteps, **kwargs)
    elif schedule == "cosine":
        sigma = torch.sqrt(1 - torch.cumprod(1 - _cosine_betas(num_timesteps, **kwargs), dim=0))
    elif schedule == "linear_sd":
        sigma = torch.sqrt(1 - torch.cumprod(1 - _linear_sd_betas(num_timesteps, **kwargs), dim=0))
    else:
        raise NotImplementedError(f"vgen_b200: sigma schedule '{schedule}' is not used by the supported configs")
    if zero_terminal_snr and abs(sigma.max() - 1.0) > 0.0001:
        sigma = _rescale_zero_terminal_snr(sigma)
    return sigma


def _f32c(x):
    return x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()


def _f16c(x):
    return x if (x.dtype == torch.float16 and x.is_contiguous()) else x.to(torch.float16).contiguous()


class GaussianDiffusion(object):
    def __init__(self, sigmas, prediction_type="eps"):
        assert prediction_type in {"x0", "eps", "v"}
        self.sigmas = sigmas.float()
        self.alphas = torch.sqrt(1 - sigmas ** 2).float()
        self.num_timesteps = len(sigmas)
        self.prediction_type = prediction_type

    # ---- host: sigma <-> t ------------------------------------------------------------------------
    def _log_sigmas(self):
        return torch.sqrt(self.sigmas ** 2 / (1 - self.sigmas ** 2)).log()

    def _sigma_to_t(self, sigma):
        """:436-456."""
        sigma = torch.as_tensor(sigma, dtype=torch.float32)
        if sigma == float("inf"):
            t = torch.full_like(sigma, len(self.sigmas) - 1)
        else:
            ls = self._log_sigmas()
            d = sigma.log() - ls[:, None]
            lo = d.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=ls.shape[0] - 2)
            hi = lo + 1
            w = ((ls[lo] - sigma.log()) / (ls[lo] - ls[hi])).clamp(0, 1)
            t = ((1 - w) * lo + w * hi).view(sigma.shape)
        return t.unsqueeze(0) if t.ndim == 0 else t

    def _t_to_sigma(self, t):
        """:458-464."""
        t = t.float()
        lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
        ls = self._log_sigmas()
        v = (1 - w) * ls[lo] + w * ls[hi]
        v[torch.isnan(v) | torch.isinf(v)] = float("inf")
        return v.exp()

    @staticmethod
    def _uniform_step(t):
        tt = t.reshape(-1)
        step = int(tt[0])
        if tt.numel() > 1 and not bool((tt == step).all()):
            raise NotImplementedError("vgen_b200 GaussianDiffusion: all batch entries must share the timestep")
        return step

    # ---- one model evaluation -> x0 ---------------------------------------------------------------
    def _predict_x0(self, xt, step, t, model, model_kwargs, guide_scale, guide_rescale):
        """model call(s) + CFG + guidance rescale + x0 (:196-230) for the table row `step`."""
        stats = None
        if guide_scale is None:
            assert isinstance(model_kwargs, dict)
            out = _f16c(model(xt, t=t, **model_kwargs))
        else:
            assert isinstance(model_kwargs, list) and len(model_kwargs) == 2
            if guide_scale == 1.0:
                out = _f16c(model(xt, t=t, **model_kwargs[0]))
            else:
                y_out, u_out = cfg_forward(model, xt, t, model_kwargs, by_keyword=True)
                out, stats = ops.cfg_combine(_f16c(y_out), _f16c(u_out), guide_scale)
                if guide_rescale is None:
                    stats = None
                else:
                    assert 0 <= guide_rescale <= 1
        return ops.gauss_x0(xt, out, float(self.alphas[step]), float(self.sigmas[step]), self.prediction_type, stats,
                            guide_rescale or 0.0)

    def _eps(self, xt, x0, step):
        a, s = float(self.alphas[step]), float(self.sigmas[step])
        return ops.lincomb_f32([(1.0 / s, xt), (-a / s, x0)])          # (xt - alphas*x0) / sigmas  (:244)

    @torch.no_grad()
    def denoise(self, xt, t, s, model, model_kwargs={}, guide_scale=None, guide_rescale=None, clamp=None, percentile=None):
        """:163-247 -> (mu, var, log_var, x0, eps) of q(x_s | x_t, x0_hat)."""
        if clamp is not None or percentile is not None:
            raise NotImplementedError("vgen_b200 GaussianDiffusion: clamp / percentile are not used on the sampling path")
        xt = _f32c(xt)
        step = self._uniform_step(t)
        sstep = step - 1 if s is None else self._uniform_step(s)
        sig, alp = self.sigmas[step], self.alphas[step]
        alp_s = self.alphas[max(sstep, 0)] if sstep >= 0 else torch.tensor(1.0)
        sig_s = torch.sqrt(1 - alp_s ** 2)
        betas = 1 - (alp / alp_s) ** 2
        coef1 = betas * alp_s / sig ** 2
        coef2 = (alp * sig_s ** 2) / (alp_s * sig ** 2)
        var = betas * (sig_s / sig) ** 2
        log_var = torch.log(var).clamp(-20, 20)
        x0 = self._predict_x0(xt, step, t, model, model_kwargs, guide_scale, guide_rescale)
        eps = self._eps(xt, x0, step)
        mu = ops.lincomb_f32([(float(coef1), x0), (float(coef2), xt)])
        shape = (xt.size(0),) + (1,) * (xt.ndim - 1)
        dev = xt.device
        return mu, var.to(dev).expand(shape), log_var.to(dev).expand(shape), x0, eps

    # ---- DDIM inversion ---------------------------------------------------------------------------
    @torch.no_grad()
    def ddim_reverse_sample(self, xt, t, model, model_kwargs={}, clamp=None, percentile=None, guide_scale=None,
                            guide_rescale=None, ddim_timesteps=20, reverse_steps=600):
        """:376-411 -> (x_{t+stride}, x0)."""
        if clamp is not None or percentile is not None:
            raise NotImplementedError("vgen_b200 GaussianDiffusion: clamp / percentile are not used on the sampling path")
        return self._ddim_reverse_step(_f32c(xt), self._uniform_step(t), t, model, model_kwargs, guide_scale, guide_rescale,
                                       ddim_timesteps, reverse_steps)

    def _ddim_reverse_step(self, xt, step, t, model, model_kwargs, guide_scale, guide_rescale, ddim_timesteps, reverse_steps):
        """`step` is the host copy of the (batch-uniform) timestep: the loop never reads `t` back from the device."""
        stride = reverse_steps // ddim_timesteps
        x0 = self._predict_x0(xt, step, t, model, model_kwargs, guide_scale, guide_rescale)
        s = min(max(step + stride, 0), reverse_steps - 1)
        a, sg = float(self.alphas[step]), float(self.sigmas[step])
        a_s = self.alphas[s]
        sg_s = float(torch.sqrt(1 - a_s ** 2))
        # mu = alphas_s*x0 + sigmas_s*eps with eps = (xt - alphas*x0)/sigmas, evaluated in one pass
        mu = ops.lincomb_f32([(float(a_s) - sg_s * a / sg, x0), (sg_s / sg, xt)])
        return mu, x0

    @torch.no_grad()
    def ddim_reverse_sample_loop(self, x0, model, model_kwargs={}, clamp=None, percentile=None, guide_scale=None,
                                 guide_rescale=None, ddim_timesteps=20, reverse_steps=600):
        """:413-434."""
        if clamp is not None or percentile is not None:
            raise NotImplementedError("vgen_b200 GaussianDiffusion: clamp / percentile are not used on the sampling path")
        b = x0.size(0)
        xt = _f32c(x0)
        for step in torch.arange(0, reverse_steps, reverse_steps // ddim_timesteps):
            t = torch.full((b,), int(step), dtype=torch.long, device=xt.device)
            xt, _ = self._ddim_reverse_step(xt, int(step), t, model, model_kwargs, guide_scale, guide_rescale, ddim_timesteps,
                                            reverse_steps)
        return xt

    # ---- DPM-Solver++(2M) SDE ---------------------------------------------------------------------
    def _sigma_ladder(self, steps, t_max, t_min, discretization, discard_penultimate_step):
        """:318-357 (non-karras)."""
        if isinstance(steps, int):
            steps += 1 if discard_penultimate_step else 0
            t_max = self.num_timesteps - 1 if t_max is None else t_max
            t_min = 0 if t_min is None else t_min
            if discretization == "leading":
                steps = torch.arange(t_min, t_max + 1, (t_max - t_min + 1) / steps).flip(0)
            elif discretization == "linspace":
                steps = torch.linspace(t_max, t_min, steps)
            elif discretization == "trailing":
                steps = torch.arange(t_max, t_min - 1, -((t_max - t_min + 1) / steps))
            else:
                raise NotImplementedError(f"{discretization} discretization not implemented")
            steps = steps.clamp_(t_min, t_max)
        steps = torch.as_tensor(steps, dtype=torch.float32, device="cpu")
        sigmas = self._t_to_sigma(steps)
        sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
        if discard_penultimate_step:
            sigmas = torch.cat([sigmas[:-2], sigmas[-1:]])
        return sigmas

    @torch.no_grad()
    def sample(self, noise, model, model_kwargs={}, condition_fn=None, guide_scale=None, guide_rescale=None, clamp=None,
               percentile=None, solver="euler_a", steps=20, t_max=None, t_min=None, discretization=None,
               discard_penultimate_step=None, return_intermediate=None, show_progress=False, seed=-1, eta=1.0,
               s_noise=1.0, solver_type="midpoint", **kwargs):
        """:250-373 with solver_fn = sample_dpmpp_2m_sde (:86-142), the only solver the reference keeps."""
        if solver != "dpmpp_2m_sde":
            raise NotImplementedError(f"vgen_b200 GaussianDiffusion: solver '{solver}' (the reference only wires dpmpp_2m_sde)")
        if clamp is not None or percentile is not None or condition_fn is not None:
            raise NotImplementedError("vgen_b200 GaussianDiffusion: clamp / percentile / condition_fn are not supported")
        assert isinstance(steps, (int, torch.LongTensor))
        assert t_max is None or (0 < t_max <= self.num_timesteps - 1)
        assert t_min is None or (0 <= t_min < self.num_timesteps - 1)
        assert discretization in (None, "leading", "linspace", "trailing")
        assert return_intermediate in (None, "x0", "xt")
        assert solver_type in {"heun", "midpoint"}
        discretization = discretization or "linspace"
        seed = seed if seed >= 0 else random.randint(0, 2 ** 31)  # drawn like upstream (:291); unused by this solver
        if isinstance(steps, torch.LongTensor):
            discard_penultimate_step = False
        if discard_penultimate_step is None:
            discard_penultimate_step = True
        sigmas = self._sigma_ladder(steps, t_max, t_min, discretization, discard_penultimate_step)

        noise = _f32c(noise)
        b = noise.size(0)
        intermediates = []

        def denoised_fn(xt, sigma):
            step = int(self._sigma_to_t(sigma).round().long()[0])
            t = torch.full((b,), step, dtype=torch.long, device=xt.device)
            x0 = self._predict_x0(xt, step, t, model, model_kwargs, guide_scale, guide_rescale)
            if return_intermediate == "xt":
                intermediates.append(xt)
            elif return_intermediate == "x0":
                intermediates.append(x0)
            return x0

        x = ops.lincomb_f32([(float(sigmas[0]), noise)])
        s_min, s_max = sigmas[sigmas > 0].min(), sigmas[sigmas < float("inf")].max()
        tree_seed = torch.randint(0, 2 ** 63 - 1, []).item()          # BatchedBrownianTree, :29-30
        tree = (BROWNIAN_TREE or brownian.default_tree_cls())(s_min, torch.zeros_like(x), s_max, entropy=tree_seed)

        def brownian_increment(s0, s1):                              # BrownianTreeNoiseSampler.__call__, :73-76
            lo, hi, sign = (s0, s1, 1.0) if s0 < s1 else (s1, s0, -1.0)
            return tree(lo, hi), sign / float((s1 - s0).abs().sqrt())

        old, h_last, h = None, None, None
        for i in range(len(sigmas) - 1):
            sg, sg_next = sigmas[i], sigmas[i + 1]
            if sg == float("inf"):
                den = denoised_fn(noise, sg)
                x = ops.lincomb_f32([(1.0, den), (float(sg_next), noise)])
            else:
                c_in = 1 / (sg ** 2 + 1.0) ** 0.5
                den = denoised_fn(ops.lincomb_f32([(float(c_in), x)]), sg)
                if sg_next == 0:
                    x = den
                else:
                    t_, s_ = -sg.log(), -sg_next.log()
                    h = s_ - t_
                    eh = eta * h
                    cx = float(sg_next / sg * (-eh).exp())
                    cd = float((-h - eh).expm1().neg())
                    w, wscale = brownian_increment(sg, sg_next)
                    cn = wscale * float(sg_next * (-2 * eh).expm1().neg().sqrt() * s_noise)
                    if old is not None:
                        r = h_last / h
                        if solver_type == "heun":
                            c2 = float(((-h - eh).expm1().neg() / (-h - eh) + 1) * (1 / r))
                        else:
                            c2 = float(0.5 * (-h - eh).expm1().neg() * (1 / r))
                        # two launches keep the reference's evaluation order (x, +2M correction, +noise)
                        x = ops.lincomb_f32([(cx, x), (cd, den)])
                        x = ops.lincomb_f32([(1.0, x), (c2, den), (-c2, old), (cn, _f32c(w))])
                    else:
                        x = ops.lincomb_f32([(cx, x), (cd, den), (cn, _f32c(w))])
            old = den
            if h is not None:
                h_last = h
        return (x, intermediates) if return_intermediate is not None else x

    @torch.no_grad()
    def diffuse(self, x0, t, noise=None):
        """:154-161."""
        x0 = _f32c(x0)
        noise = torch.randn_like(x0) if noise is None else _f32c(noise)
        step = self._uniform_step(t)
        return ops.lincomb_f32([(float(self.alphas[step]), x0), (float(self.sigmas[step]), noise)])


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class DiffusionDDIMSR(object):
    """diffusion_ddim.py:19-25 (registered as DIFFUSION 'DiffusionDDIMSR')."""

    def __init__(self, reverse_diffusion, forward_diffusion, **kwargs):
        self.reverse_diffusion = GaussianDiffusion(
            sigmas=sigma_schedule(_cfg_get(reverse_diffusion, "schedule"), **_cfg_get(reverse_diffusion, "schedule_param")),
            prediction_type=_cfg_get(reverse_diffusion, "mean_type"))
        self.forward_diffusion = GaussianDiffusion(
            sigmas=sigma_schedule(_cfg_get(forward_diffusion, "schedule"), **_cfg_get(forward_diffusion, "schedule_param")),
            prediction_type=_cfg_get(forward_diffusion, "mean_type"))
