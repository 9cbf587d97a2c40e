"""Run one op a few times (for ncu).  usage: prof_one.py attn|linear|conv"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vgen_b200 import ops
g = torch.Generator().manual_seed(0)
what = sys.argv[1]
if what == "attn":
    b, h, l = 16, 5, 14080
    qkv = torch.randn(b, l, 3 * h * 64, generator=g).half().cuda()
    fn = lambda: ops.attention_d64(qkv[:, :, :h * 64], qkv[:, :, h * 64:2 * h * 64], qkv[:, :, 2 * h * 64:], h)
elif what == "geglu":
    m, k, inner = 225280, 320, 1280
    a = torch.randn(m, k, generator=g).half().cuda(); w = (torch.randn(2 * inner, k, generator=g) * k ** -0.5).half().cuda()
    bias = torch.randn(2 * inner, generator=g)
    wp, bp = ops.pack_geglu_weight(w.cpu().float(), bias, 256)
    wp, bp = wp.half().cuda(), bp.cuda(); o = torch.empty(m, inner, dtype=torch.float16, device="cuda")
    fn = lambda: ops.linear(a, wp, bias=bp, geglu=True, out=o, bn=256)
elif what == "linear":
    m, k, n = 225280, 320, 320
    a = torch.randn(m, k, generator=g).half().cuda(); w = (torch.randn(n, k, generator=g) * k ** -0.5).half().cuda()
    bias = torch.randn(n, generator=g).cuda(); r = torch.randn(m, n, generator=g).half().cuda(); o = torch.empty(m, n, dtype=torch.float16, device="cuda")
    fn = lambda: ops.linear(a, w, bias=bias, residual=r, out=o)
else:
    x = torch.randn(16, 88, 160, 320, generator=g).half().cuda(); w = (torch.randn(320, 2880, generator=g) * 2880 ** -0.5).half().cuda()
    bias = torch.randn(320, generator=g).cuda(); o = torch.empty(16, 88, 160, 320, dtype=torch.float16, device="cuda")
    fn = lambda: ops.conv2d_3x3(x, w, bias=bias, out=o)
for _ in range(3):
    fn()
torch.cuda.synchronize()
