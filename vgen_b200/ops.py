"""Tensor-level wrappers over the C ABI: they check shapes/dtypes, allocate outputs with torch
(device memory plumbing only) and pass raw pointers + the current CUDA stream to libvgen_b200.so.

Layout convention: activations fp16 channels-last ("rows x channels"); weights fp16 [out][taps*in];
bias / affine parameters fp32.
"""
from __future__ import annotations

import ctypes

import torch

from . import lib as _l
from .lib import Epilogue


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk16(t, name):
    if t.dtype != torch.float16 or not t.is_cuda:
        raise _l.VgenError(f"{name}: expected a CUDA fp16 tensor, got {t.dtype} on {t.device}")


def _rows_view(t, name):
    """(rows, ld) of a tensor whose last dim is contiguous and whose leading dims are row-dense."""
    if t.stride(-1) != 1:
        raise _l.VgenError(f"{name}: innermost dimension must be contiguous")
    t2 = t.reshape(-1, t.shape[-1]) if t.is_contiguous() else t
    if t2.dim() != 2:
        raise _l.VgenError(f"{name}: non-contiguous tensors must be 2-D row views")
    return t2, t2.stride(0)


def _epilogue(n_out, bias=None, group_bias=None, residual=None, alpha=1.0, geglu=False, bn=0):
    e = Epilogue()
    e.alpha = float(alpha)
    keep = []
    if bias is not None:
        if bias.dtype != torch.float32:
            raise _l.VgenError("bias must be fp32")
        e.bias = bias.data_ptr()
        keep.append(bias)
    if group_bias is not None:
        _chk16(group_bias, "group_bias")
        e.group_bias = group_bias.data_ptr()
        e.group_bias_ld = group_bias.stride(0)
        keep.append(group_bias)
    if residual is not None:
        _chk16(residual, "residual")
        r2, ldr = _rows_view(residual, "residual")
        if r2.shape[-1] != n_out:
            raise _l.VgenError(f"residual has {r2.shape[-1]} columns, expected {n_out}")
        e.residual = r2.data_ptr()
        e.residual_ld = ldr
        keep.append(r2)
    e.geglu = 1 if geglu else 0
    e.bn = int(bn)
    return e, keep


def linear(a, w, bias=None, residual=None, alpha=1.0, geglu=False, out=None, bn=0):
    """out[m, n] = epi(a[m, k] @ w[n, k]^T).  `a` may be a row-strided 2-D view."""
    _chk16(a, "a"), _chk16(w, "w")
    lead = a.shape[:-1]
    a2, lda = _rows_view(a, "a")
    m, k = a2.shape
    n = w.shape[0]
    if w.shape[1] != k or not w.is_contiguous():
        raise _l.VgenError(f"linear: weight {tuple(w.shape)} does not match k={k}")
    n_out = n // 2 if geglu else n
    if out is None:
        out = torch.empty(*lead, n_out, device=a.device, dtype=torch.float16)
    o2, ldo = _rows_view(out, "out")
    e, keep = _epilogue(n_out, bias, None, residual, alpha, geglu, bn)
    rc = _l.load().vgen_linear(_p(a2), m, k, lda, _p(w), n, _p(o2), ldo, ctypes.byref(e), _stream())
    _l.check(rc, "vgen_linear")
    return out


def conv2d_3x3(x, w, bias=None, group_bias=None, residual=None, out=None, bn=0):
    """x [nimg, h, w, c] fp16 -> [nimg, h, w, n]; w [n, 9*c] with k = (ky*3+kx)*c + ci."""
    _chk16(x, "x"), _chk16(w, "w")
    if not x.is_contiguous() or x.dim() != 4:
        raise _l.VgenError("conv2d_3x3: x must be contiguous [nimg,h,w,c]")
    nimg, h, wd, c = x.shape
    n = w.shape[0]
    if w.shape[1] != 9 * c or not w.is_contiguous():
        raise _l.VgenError(f"conv2d_3x3: weight {tuple(w.shape)} does not match c={c}")
    if out is None:
        out = torch.empty(nimg, h, wd, n, device=x.device, dtype=torch.float16)
    o2, ldo = _rows_view(out, "out")
    e, keep = _epilogue(n, bias, group_bias, residual, 1.0, False, bn)
    rc = _l.load().vgen_conv2d_3x3(_p(x), nimg, h, wd, c, _p(w), n, _p(o2), ldo, ctypes.byref(e), _stream())
    _l.check(rc, "vgen_conv2d_3x3")
    return out


def tconv3(x, w, bias=None, residual=None, out=None, bn=0):
    """x [f, hw, c] fp16 -> [f, hw, n]; temporal 3-tap conv, zero padded; w [n, 3*c], k = kt*c + ci."""
    _chk16(x, "x"), _chk16(w, "w")
    if not x.is_contiguous() or x.dim() != 3:
        raise _l.VgenError("tconv3: x must be contiguous [f,hw,c]")
    f, hw, c = x.shape
    n = w.shape[0]
    if w.shape[1] != 3 * c or not w.is_contiguous():
        raise _l.VgenError(f"tconv3: weight {tuple(w.shape)} does not match c={c}")
    if out is None:
        out = torch.empty(f, hw, n, device=x.device, dtype=torch.float16)
    o2, ldo = _rows_view(out, "out")
    e, keep = _epilogue(n, bias, None, residual, 1.0, False, bn)
    rc = _l.load().vgen_tconv3(_p(x), f, hw, c, _p(w), n, _p(o2), ldo, ctypes.byref(e), _stream())
    _l.check(rc, "vgen_tconv3")
    return out


def set_tapgemm_impl(impl: str):
    _l.check(_l.load().vgen_set_tapgemm_impl({"sm100": 0, "simt": 1}[impl]), "vgen_set_tapgemm_impl")


def pack_geglu_weight(w, bias, bn):
    """Interleave GEGLU.proj rows so every bn-block holds bn/2 value rows then their bn/2 gate rows.

    w: [2*inner, k] (first half = value, second half = gate, reference util.py:711-714).
    """
    two_inner, k = w.shape
    inner = two_inner // 2
    hb = bn // 2
    assert inner % hb == 0
    v = w[:inner].reshape(inner // hb, hb, k)
    g = w[inner:].reshape(inner // hb, hb, k)
    wp = torch.cat([v, g], dim=1).reshape(two_inner, k).contiguous()
    bp = None
    if bias is not None:
        bv = bias[:inner].reshape(inner // hb, hb)
        bg = bias[inner:].reshape(inner // hb, hb)
        bp = torch.cat([bv, bg], dim=1).reshape(two_inner).contiguous()
    return wp, bp
