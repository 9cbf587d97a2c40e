"""TEST INFRASTRUCTURE ONLY -- golden vectors for the CLIP conditioning path from the reference's own code:
the vendored open_clip `CLIP` class (utils/reward/open_clip/model.py) driven exactly like
tools/modules/clip_embedder.py:183-212 drives it, on a tiny configuration with synthetic weights, plus tokenizer ids of
fixed prompts from the vendored tokenizer.  Run in the build container:  python -m oracle.make_golden_clip
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np
import torch

from . import clip_oracle as co, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = "/root/reference"
TINY = dict(embed_dim=64, vision_cfg=dict(image_size=56, layers=2, width=160, head_width=80, patch_size=14),
            text_cfg=dict(context_length=77, vocab_size=49408, width=128, heads=2, layers=3))
PROMPTS = ["", "a photo of a cat", "An astronaut riding a horse on Mars, 4k, highly detailed!!",
           "Pikachu turn back & wave... it's <b>bold</b> &amp; naive cafe No5 2024-05-01",
           "Distorted, discontinuous, Ugly, blurry, low resolution, motionless, static, disfigured, disconnected limbs, Ugly faces, incomplete arms",
           " ".join(["supercalifragilisticexpialidocious"] * 30), "don't we'll they've I'm he'd"]


def load_reference_open_clip():
    """The vendored package's __init__ pulls in unrelated modules (turtle -> tkinter): register empty parents instead."""
    if R not in sys.path:
        sys.path.insert(0, R)
    if "ftfy" not in sys.modules:
        ft = types.ModuleType("ftfy")
        ft.fix_text = lambda t: t
        sys.modules["ftfy"] = ft
    for pkg, path in [("utils.reward", R + "/utils/reward"), ("utils.reward.open_clip", R + "/utils/reward/open_clip")]:
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [path]
            sys.modules[pkg] = m
    return importlib.import_module("utils.reward.open_clip.model"), importlib.import_module("utils.reward.open_clip.tokenizer")


def clip_state_dict(model_mod, cfg, seed):
    m = model_mod.CLIP(**cfg)
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = synth.state_dict([s for s in spec if len(s[1]) > 0], seed=seed)
    for k in ("positional_embedding", "visual.positional_embedding", "visual.class_embedding"):
        sd[k] = synth.tensor(k, dict(spec)[k], 0.02, seed)           # open_clip's own scale for the embeddings
    sd["token_embedding.weight"] = synth.tensor("token_embedding.weight", dict(spec)["token_embedding.weight"], 0.02, seed)
    sd["logit_scale"] = torch.tensor(2.6592)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd, spec


def reference_forward(m, tokens, image, layer_idx):
    """clip_embedder.py:183-212 on the open_clip model `m`."""
    with torch.no_grad():
        xi = m.encode_image(image)
        x = m.token_embedding(tokens) + m.positional_embedding
        x = x.permute(1, 0, 2)
        for i, r in enumerate(m.transformer.resblocks):
            if i == len(m.transformer.resblocks) - layer_idx:
                break
            x = r(x, attn_mask=m.attn_mask)
        x = m.ln_final(x.permute(1, 0, 2))
        xt = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ m.text_projection
    return xi, xt, x


def main():
    model_mod, tok_mod = load_reference_open_clip()
    m, sd, spec = clip_state_dict(model_mod, TINY, seed=31)
    tokens = tok_mod.tokenize(PROMPTS)
    image = synth.tensor("clip_image", (2, 3, 56, 56), 1.0, 31)
    out = {"tokens": tokens.numpy()}
    rep = {}
    for name, li in (("last", 0), ("penultimate", 1)):
        xi, xt, x = reference_forward(m, tokens, image, li)
        oxt, ox = co.encode_text(sd, tokens, TINY["text_cfg"]["heads"], name)
        oxi = co.encode_image(sd, image, TINY["vision_cfg"]["head_width"])
        rep[name] = {"xt": float((oxt - xt).abs().max() / xt.abs().max()), "x": float((ox - x).abs().max() / x.abs().max()),
                     "xi": float((oxi - xi).abs().max() / xi.abs().max())}
        out[f"{name}_xt"], out[f"{name}_x"], out["xi"] = xt.numpy(), x.numpy(), xi.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clip_tiny.npz"), **out)
    json.dump([[k, list(s)] for k, s in spec], open(os.path.join(ROOT, "tests", "golden", "clip_tiny.spec.json"), "w"))
    print("oracle vs vendored open_clip (max-rel):", rep)


if __name__ == "__main__":
    main()
