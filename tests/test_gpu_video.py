"""Video write-out on the B200: the device kernel behind vgen_b200.video_io against the frames
the REAL reference function produced (tests/golden/video_out.npz) and against the numpy oracle at the BASELINE size.
Byte work: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_video as mg, video_oracle as vo
from vgen_b200 import ops, video_io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(mg.CASES))
def test_frames_match_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "video_out.npz"))
    v, mean, std = mg.make_video(name)
    host, band, ev = video_io.frames_to_host(v.cuda(), mean, std)
    ev.synchronize()
    n = video_io.select_frames(host, band)
    gold = g[name + "_frames"]
    assert n == gold.shape[0]
    assert np.array_equal(host[:n].numpy(), gold)
    # the anomaly counters themselves
    assert [int(b) for b in band] == vo.band_counts(vo.frames_uint8(v.numpy(), mean, std))


@pytest.mark.parametrize("shape", [(1, 3, 16, 704, 1280), (1, 3, 2, 37, 53), (1, 3, 1, 8, 8)])
def test_frames_match_oracle_at_size(shape):
    g = torch.Generator().manual_seed(3)
    v = torch.randn(shape, generator=g) * 0.8
    v[0, :, -1, : shape[3] // 2] = 0.0                              # a half-grey last frame
    out, band = ops.video_to_rgb8(v[0].cuda().contiguous(), [0.5, 0.5, 0.5], [0.5, 0.5, 0.5])
    ref = np.stack(vo.frames_uint8(v.numpy()))
    assert np.array_equal(out.cpu().numpy(), ref)
    assert [int(b) for b in band.cpu()] == vo.band_counts(list(ref))


def test_save_video_roundtrip(tmp_path):
    """The drop-in writes a file through whichever encoder exists (ffmpeg pipe, else OpenCV); async == sync bytes in."""
    import shutil
    try:
        import cv2  # noqa: F401
        have = True
    except ImportError:
        have = bool(shutil.which("ffmpeg"))
    if not have:
        pytest.skip("no encoder (ffmpeg / OpenCV) on this box")
    v = (torch.rand(1, 3, 8, 64, 96) * 2 - 1).cuda()
    p = video_io.save_i2vgen_video_safe(str(tmp_path / "a.mp4"), v, ["x"], save_fps=8)
    assert os.path.exists(p) and os.path.getsize(p) > 0
    h = video_io.save_t2vhigen_video_safe(str(tmp_path / "b.mp4"), v, ["x"], async_=True)
    p2 = h.join()
    assert os.path.exists(p2) and os.path.getsize(p2) > 0
    p3 = video_io.save_i2vgen_video_safe(str(tmp_path / "c"), v[:, :, :1], ["x"])
    assert p3.endswith(".png") and os.path.exists(p3)
