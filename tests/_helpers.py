"""Shared test plumbing: build the vgen_b200 class of a parity case, and call product / oracle alike."""
from __future__ import annotations

import vgen_b200
from oracle import vgen_oracle as vo
from oracle.cases import LCM_CONFIG

_CLASSES = {"t2v": "UNetSD_T2VBase", "i2vgen": "UNetSD_I2VGen", "videolcm": "UNetSD_VideoLCM", "sr600": "UNetSD_SR600",
            "higen": "UNetSD_HiGen", "vae": "AutoencoderKL"}


def product_class(kind):
    return getattr(vgen_b200, _CLASSES[kind])


def build_product(case):
    cls = product_class(case["kind"])
    if case["kind"] == "videolcm":
        return cls(config=dict(LCM_CONFIG), **case["ctor"])
    return cls(**case["ctor"])


def _higen_kw(inp):
    return dict(spat_prior=inp["spat_prior"], motion_cond=inp["motion_cond"], appearance_cond=inp["appearance_cond"])


def product_call(case, m, inp):
    """The call an inference entrance makes on the MODEL (same keyword names as the reference)."""
    k = case["kind"]
    if k in ("t2v", "videolcm"):
        return m(inp["x"], inp["t"], y=inp["y"])
    if k == "sr600":
        return m(inp["x"], inp["t"], inp["y"])
    if k == "i2vgen":
        return m(inp["x"], inp["t"], y=inp["y"], image=inp["image"], local_image=inp["local_image"], fps=inp["fps"])
    if k == "higen":
        return m(inp["x"], inp["t"], y=inp["y"], **_higen_kw(inp))
    return m.decode(inp["z"])


def oracle_call(case, sd, inp):
    k = case["kind"]
    if k == "t2v":
        return vo.unet_t2v_forward(sd, inp["x"], inp["t"], inp["y"])
    if k == "videolcm":
        return vo.unet_videolcm_forward(sd, inp["x"], inp["t"], inp["y"])
    if k == "sr600":
        return vo.unet_sr600_forward(sd, inp["x"], inp["t"], inp["y"])
    if k == "i2vgen":
        return vo.unet_i2vgen_forward(sd, inp["x"], inp["t"], inp["y"], inp["image"], inp["local_image"], inp["fps"])
    if k == "higen":
        return vo.unet_higen_forward(sd, inp["x"], inp["t"], inp["y"], **_higen_kw(inp))
    return vo.vae_decode(sd, inp["z"])
