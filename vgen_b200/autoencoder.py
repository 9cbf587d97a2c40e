"""B200-native AutoencoderKL (SD-2.1 VAE): same constructor, state_dict keys and methods as the
reference class (tools/modules/autoencoder.py:30-103); `decode` is a channels-last graph of
libvgen_b200.so kernels (tcgen05 conv3x3 / 1x1, fused GroupNorm+SiLU, GEMM-based mid attention).

The reference runs the decoder in fp32 (outside its autocast block, inference_i2vgen_entrance.py:224-230);
here activations are fp16 with fp32 accumulation and fp32 norm/softmax statistics, and the result is
returned as fp32 like the reference.  `encode_firsr_stage` (conditioning side)
runs the Encoder (:483-578) on the same kernels and draws the posterior noise with the CPU generator exactly
like the reference (:223-225) so RNG streams stay aligned.
"""
from __future__ import annotations

import collections

import torch

from . import arch, ops
from .graph import graphed
from .params import SpecModule
from .unet import _f16, _f32, _pack_conv3x3


class AutoencoderKL(SpecModule):
    def __init__(self, ddconfig, embed_dim, pretrained=None, ignore_keys=[], image_key="image", colorize_nlabels=None,
                 monitor=None, ema_decay=None, learn_logvar=False, use_vid_decoder=False, **kwargs):
        super().__init__()
        assert ddconfig["double_z"]
        self.plan = arch.vae_plan(dict(ddconfig), embed_dim)
        self.embed_dim = embed_dim
        self.image_key = image_key
        self.learn_logvar = learn_logvar
        self._build_params(arch.vae_spec(self.plan))
        if pretrained is not None:
            self.init_from_ckpt(pretrained, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=list()):
        """autoencoder.py:64-73: keys containing 'first_stage_model.' of an SD checkpoint, strict."""
        sd = torch.load(path, map_location="cpu")["state_dict"]
        new = collections.OrderedDict()
        for k in list(sd.keys()):
            if k.find("first_stage_model") >= 0:
                new[k.split("first_stage_model.")[-1]] = sd[k]
        self.load_state_dict(new, strict=True)

    # ------------------------------------------------------------------------------ packing
    def _pack(self):
        sd = self.state_dict()
        dev = self.device
        if dev.type != "cuda":
            raise ops._l.VgenError("vgen_b200 AutoencoderKL.decode needs the module on a CUDA device (no CPU path exists)")
        W = {}

        def norm(p):
            W[p + "g"], W[p + "b"] = _f32(sd[p + "weight"], dev), _f32(sd[p + "bias"], dev)

        def conv3(p, cin_pad=None):
            W[p + "w"], W[p + "b"] = _pack_conv3x3(sd[p + "weight"], dev, cin_pad), _f32(sd[p + "bias"], dev)

        def conv1(p):
            w = sd[p + "weight"]
            W[p + "w"], W[p + "b"] = _f16(w.reshape(w.shape[0], -1), dev), _f32(sd[p + "bias"], dev)

        def fuse_qkv(a):   # one [3c, c] projection for the flash-attention path
            W[a + "qkv.w"] = torch.cat([W[a + "q.w"], W[a + "k.w"], W[a + "v.w"]], 0).contiguous()
            W[a + "qkv.b"] = torch.cat([W[a + "q.b"], W[a + "k.b"], W[a + "v.b"]], 0).contiguous()

        def resnet(p):
            norm(p + "norm1."), conv3(p + "conv1."), norm(p + "norm2."), conv3(p + "conv2.")
            if (p + "nin_shortcut.weight") in sd:
                conv1(p + "nin_shortcut.")

        conv1("post_quant_conv.")
        conv3("decoder.conv_in.", cin_pad=8)
        resnet("decoder.mid.block_1."), resnet("decoder.mid.block_2.")
        a = "decoder.mid.attn_1."
        norm(a + "norm.")
        for nm in ("q.", "k.", "v.", "proj_out."):
            conv1(a + nm)
        fuse_qkv(a)
        for lvl in range(len(self.plan.ch_mult)):
            for j in range(self.plan.num_res_blocks + 1):
                resnet(f"decoder.up.{lvl}.block.{j}.")
            if lvl != 0:
                conv3(f"decoder.up.{lvl}.upsample.conv.")
        norm("decoder.norm_out."), conv3("decoder.conv_out.")
        # encoder (autoencoder.py:483-547)
        conv3("encoder.conv_in.", cin_pad=8)
        nres = len(self.plan.ch_mult)
        for lvl in range(nres):
            for j in range(self.plan.num_res_blocks):
                resnet(f"encoder.down.{lvl}.block.{j}.")
            if lvl != nres - 1:
                conv3(f"encoder.down.{lvl}.downsample.conv.")
        resnet("encoder.mid.block_1."), resnet("encoder.mid.block_2.")
        a = "encoder.mid.attn_1."
        norm(a + "norm.")
        for nm in ("q.", "k.", "v.", "proj_out."):
            conv1(a + nm)
        fuse_qkv(a)
        norm("encoder.norm_out."), conv3("encoder.conv_out."), conv1("quant_conv.")
        self._packed = W
        return W

    # ------------------------------------------------------------------------------ blocks
    @staticmethod
    def _conv3(x, W, p, residual=None):
        n, h, w, c = x.shape
        wt = W[p + "w"]
        if c % 64 == 0 and wt.shape[1] == 9 * c:
            return ops.conv2d_3x3(x, wt, bias=W[p + "b"], residual=residual)
        col = ops.im2col(x, 3, 3, 1, 1, 1, h, w, wt.shape[1])
        return ops.linear(col, wt, bias=W[p + "b"], residual=residual).view(n, h, w, wt.shape[0])

    def _resnet(self, x, W, p):
        """ResnetBlock.forward, autoencoder.py:315-335 (temb None)."""
        n, h, w, cin = x.shape
        g = ops.group_norm(x, W[p + "norm1.g"], W[p + "norm1.b"], 1e-6, True)
        hcur = self._conv3(g, W, p + "conv1.")
        g = ops.group_norm(hcur, W[p + "norm2.g"], W[p + "norm2.b"], 1e-6, True)
        if (p + "nin_shortcut.w") in W:
            skip = ops.linear(x.view(-1, cin), W[p + "nin_shortcut.w"], bias=W[p + "nin_shortcut.b"])
        else:
            skip = x.view(-1, cin)
        return self._conv3(g, W, p + "conv2.", residual=skip)

    def _attn(self, x, W, p):
        """AttnBlock.forward, autoencoder.py:365-389: single-head attention over h*w tokens of width c.
        c = 512 (the SD VAE) / c = 64: tcgen05 flash attention on a fused q|k|v projection -- the [hw, hw] score matrix
        (793 MB fp32 per 1280x704 frame in the reference) never leaves the SM, scores / softmax statistics are fp32.
        Other widths: three GEMMs + a row softmax per image (S materialised in fp16)."""
        n, h, w, c = x.shape
        hw = h * w
        g = ops.group_norm(x, W[p + "norm.g"], W[p + "norm.b"], 1e-6, False).view(n, hw, c)
        if c in (64, 512):
            qkv = ops.linear(g, W[p + "qkv.w"], bias=W[p + "qkv.b"])                # [n, hw, 3c]
            q, k, v = qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:]
            o = ops.attention_d512(q, k, v) if c == 512 else ops.attention_d64(q, k, v, 1)
        else:
            q = ops.linear(g, W[p + "q.w"], bias=W[p + "q.b"])
            k = ops.linear(g, W[p + "k.w"], bias=W[p + "k.b"])
            o = torch.empty(n, hw, c, device=x.device, dtype=torch.float16)
            for i in range(n):
                # v's bias is added after P v (rows of P sum to 1), so v^T can be produced directly as W_v g^T
                vt = ops.linear(W[p + "v.w"], g[i])                                # [c, hw] = W_v g^T
                s = ops.linear(q[i], k[i], alpha=float(c) ** -0.5)                 # [hw, hw]
                ops.softmax_rows_(s, 1.0)
                ops.linear(s, vt, bias=W[p + "v.b"], out=o[i])                     # [hw, c]
        out = ops.linear(o.view(-1, c), W[p + "proj_out.w"], bias=W[p + "proj_out.b"], residual=x.view(-1, c))
        return out.view(n, h, w, c)

    # ------------------------------------------------------------------------------ public API
    @graphed
    @torch.no_grad()
    def decode(self, z, **kwargs):
        """autoencoder.py:100-103 -> Decoder.forward :653-686.  z [n, 4, h, w] -> fp32 [n, 3, 8h, 8w]."""
        if not z.is_cuda:
            raise ops._l.VgenError("vgen_b200 AutoencoderKL.decode: z must be a CUDA tensor (no CPU path exists)")
        W = self._packed or self._pack()
        n, zc, h, w = z.shape
        x = ops.cp_to_pc(z.contiguous(), n, zc, h * w)                                    # [n, hw, zc]
        x = ops.linear_small(x.view(-1, zc), W["post_quant_conv.w"], W["post_quant_conv.b"])
        x8 = torch.zeros(n * h * w, 8, device=z.device, dtype=torch.float16)
        ops.copy2d(x, x8[:, :zc])
        d = "decoder."
        hcur = self._conv3(x8.view(n, h, w, 8), W, d + "conv_in.")
        hcur = self._resnet(hcur, W, d + "mid.block_1.")
        hcur = self._attn(hcur, W, d + "mid.attn_1.")
        hcur = self._resnet(hcur, W, d + "mid.block_2.")
        for lvl in reversed(range(len(self.plan.ch_mult))):
            for j in range(self.plan.num_res_blocks + 1):
                hcur = self._resnet(hcur, W, f"{d}up.{lvl}.block.{j}.")
            if lvl != 0:
                hcur = self._conv3(ops.upsample_nearest2x(hcur), W, f"{d}up.{lvl}.upsample.conv.")
        g = ops.group_norm(hcur, W[d + "norm_out.g"], W[d + "norm_out.b"], 1e-6, True)
        out = self._conv3(g, W, d + "conv_out.")                                          # [n, H, W, 3]
        nn_, hh, ww, oc = out.shape
        return ops.pc_to_cp(out.view(nn_, hh * ww, oc), nn_, oc, hh * ww, torch.float32).view(nn_, oc, hh, ww)

    @torch.no_grad()
    def _encode_moments(self, x):
        """Encoder.forward (autoencoder.py:549-578) + quant_conv (:87): x [n, 3, H, W] -> moments fp16 [n, hw, 2*zc]."""
        if not x.is_cuda:
            raise ops._l.VgenError("vgen_b200 AutoencoderKL.encode: x must be a CUDA tensor (no CPU path exists)")
        W = self._packed or self._pack()
        n, c, H, Wd = x.shape
        e = "encoder."
        h = ops.cp_to_pc(x.float().contiguous(), n, c, H * Wd, c_pad=8).view(n, H, Wd, 8)
        h = self._conv3(h, W, e + "conv_in.")
        nres = len(self.plan.ch_mult)
        for lvl in range(nres):
            for j in range(self.plan.num_res_blocks):
                h = self._resnet(h, W, f"{e}down.{lvl}.block.{j}.")
            if lvl != nres - 1:
                # Downsample: F.pad (0,1,0,1) then conv 3x3 stride 2 pad 0 (autoencoder.py:462-480)
                nn_, hh, ww, cc = h.shape
                ho, wo = (hh + 1 - 3) // 2 + 1, (ww + 1 - 3) // 2 + 1
                p = f"{e}down.{lvl}.downsample.conv."
                col = ops.im2col(h, 3, 3, 2, 0, 0, ho, wo, W[p + "w"].shape[1])
                h = ops.linear(col, W[p + "w"], bias=W[p + "b"]).view(nn_, ho, wo, W[p + "w"].shape[0])
        h = self._resnet(h, W, e + "mid.block_1.")
        h = self._attn(h, W, e + "mid.attn_1.")
        h = self._resnet(h, W, e + "mid.block_2.")
        g = ops.group_norm(h, W[e + "norm_out.g"], W[e + "norm_out.b"], 1e-6, True)
        m = self._conv3(g, W, e + "conv_out.")                                            # [n, h, w, 2*zc]
        nn_, hh, ww, c2 = m.shape
        mom = ops.linear_small(m.view(-1, c2), W["quant_conv.w"], W["quant_conv.b"])
        return mom.view(nn_, hh * ww, c2), hh, ww

    @torch.no_grad()
    def encode_firsr_stage(self, x, scale_factor=1.0):
        """autoencoder.py:85-90: z = scale_factor * (mean + std * randn), noise drawn on the CPU generator (:223-225)."""
        mom, hh, ww = self._encode_moments(x)
        n, _, c2 = mom.shape
        zc = c2 // 2
        noise = torch.randn(n, zc, hh, ww).to(device=x.device)
        z = ops.vae_sample(mom, noise.view(n, zc, hh * ww).contiguous(), scale_factor)
        return z.view(n, zc, hh, ww)

    def encode(self, x):
        raise NotImplementedError("vgen_b200: AutoencoderKL.encode (returns a distribution object; training-side) is not "
                                  "provided; use encode_firsr_stage")

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("vgen_b200: AutoencoderKL.forward (training) is out of scope")
