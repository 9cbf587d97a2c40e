"""Per-op device time of one AutoencoderKL.decode call (2 frames, 1280x704) via ops.KernelProfile, plus the graph-replayed total."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vgen_b200  # noqa: E402
from vgen_b200 import ops  # noqa: E402

VAE_KW = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                            ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)
vae = vgen_b200.AutoencoderKL(**VAE_KW)
g = torch.Generator().manual_seed(99)
for p in vae.parameters():
    if p.dim() > 1:
        p.data.normal_(0, (p[0].numel()) ** -0.5, generator=g)
vae = vae.cuda().eval()
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (88, 160)
z = torch.randn(2, 4, h, w, generator=g).cuda()
vae.decode(z)
vae.decode(z)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    vae.decode(z)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print("decode(2 frames) graph replay ms:", sorted(ts))
ops.PROF = ops.KernelProfile()
vae.decode(z)
shapes = ops.PROF.by_shape()
ops.PROF = None
tot = sum(v["ms"] for v in shapes.values())
print("instrumented eager total ms", round(tot, 3))
for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:24]:
    print(f"{v['ms']:8.3f} ms  x{v['launches']:<3d} {v['flops'] / 1e9 / max(v['ms'], 1e-9):8.1f} TF  {k}")
