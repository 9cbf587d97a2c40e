"""TEST INFRASTRUCTURE ONLY -- imports the *actual* reference (ali-vilab/VGen, read-only at
/root/reference) on CPU so the restatement in oracle/vgen_oracle.py can be pinned against it and
golden vectors can be generated (oracle/make_golden.py).  Nothing here travels to the GPU box:
`available()` is False there and every user must skip.

The reference needs a GPU-only software stack (xformers, fairscale, open_clip, rotary_embedding_torch)
that is not installed; the stubs below provide the same arithmetic on CPU:
  * xformers.ops.memory_efficient_attention -> torch SDPA (exact softmax attention, scale d^-0.5;
    reference call sites tools/modules/unet/util.py:254,259)
  * fairscale checkpoint_wrapper            -> identity (it is a no-op under no_grad)
  * Tensor.cuda                             -> identity (unet_i2vgen.py:283 hard-codes .cuda())
  * torchsde.BrownianTree                   -> oracle/brownian.py (a seeded Brownian-bridge tree with the same
    call interface; torchsde is a pip dependency of the reference, absent here and on the GPU box, so
    the stochastic term of sample_dpmpp_2m_sde is pinned on the SAME stub for both sides)
and fake parent packages keep tools/__init__.py (which imports every engine) from running.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("VGEN_REFERENCE_ROOT", "/root/reference")
_loaded = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "tools", "modules", "unet"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch
    import torch.nn.functional as F

    def mea(q, k, v, attn_bias=None, op=None):
        return F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)

    if "xformers" not in sys.modules:
        ops = _stub("xformers.ops", memory_efficient_attention=mea)
        _stub("xformers", ops=ops)
    for name in ("open_clip",):
        if name not in sys.modules:
            _stub(name)
    if "rotary_embedding_torch" not in sys.modules:
        _stub("rotary_embedding_torch", RotaryEmbedding=object)
    if "fairscale" not in sys.modules:
        _stub("fairscale")
        _stub("fairscale.nn")
        _stub("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
    if "torchsde" not in sys.modules:
        from . import brownian
        _stub("torchsde", BrownianTree=brownian.BrownianTree)
    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v
        _stub("easydict", EasyDict=EasyDict)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    for pkg, sub in (("tools", "tools"), ("tools.modules", "tools/modules"),
                     ("tools.modules.unet", "tools/modules/unet"),
                     ("tools.modules.diffusions", "tools/modules/diffusions")):
        if pkg not in sys.modules:
            _stub(pkg).__path__ = [os.path.join(REF_ROOT, sub)]


def load():
    """Return a namespace with the reference classes (UNetSD_T2VBase, UNetSD_I2VGen, DiffusionDDIM,
    AutoencoderKL, schedules module).  Raises RuntimeError when the reference is not mounted."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _install_stubs()
    t2v = importlib.import_module("tools.modules.unet.unet_t2v")
    i2v = importlib.import_module("tools.modules.unet.unet_i2vgen")
    ddim = importlib.import_module("tools.modules.diffusions.diffusion_ddim")
    sched = importlib.import_module("tools.modules.diffusions.schedules")
    ae = importlib.import_module("tools.modules.autoencoder")
    util = importlib.import_module("tools.modules.unet.util")
    reg = importlib.import_module("utils.registry_class")
    lcm = importlib.import_module("tools.modules.unet.unet_videolcm")
    sr = importlib.import_module("tools.modules.unet.unet_sr600")
    hig = importlib.import_module("tools.modules.unet.unet_higen")
    gauss = importlib.import_module("tools.modules.diffusions.diffusion_gauss")
    _loaded.update(UNetSD_VideoLCM=lcm.UNetSD_VideoLCM, UNetSD_SR600=sr.UNetSD_SR600, UNetSD_HiGen=hig.UNetSD_HiGen,
                   diffusion_gauss=gauss, diffusion_ddim=ddim)
    _loaded.update(UNetSD_T2VBase=t2v.UNetSD_T2VBase, UNetSD_I2VGen=i2v.UNetSD_I2VGen,
                   DiffusionDDIM=ddim.DiffusionDDIM, schedules=sched, AutoencoderKL=ae.AutoencoderKL,
                   util=util, registry=reg)
    return types.SimpleNamespace(**_loaded)
