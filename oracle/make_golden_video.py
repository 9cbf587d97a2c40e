"""TEST INFRASTRUCTURE ONLY -- freeze what the REAL reference hands to its video encoder.

Runs utils/video_op.py:save_i2vgen_video_safe from /root/reference on small synthetic videos with `imageio` replaced by
a writer that records the frames (imageio / ffmpeg are not installed), and stores inputs' recipes + frames in
tests/golden/video_out.npz.  Run in the build container:  python -m oracle.make_golden_video
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    # name: (shape [b,3,f,h,w], scale, mean, std, grey_last)
    "plain": ((1, 3, 4, 6, 10), 1.2, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], False),
    "ragged_meanstd": ((2, 3, 3, 5, 7), 2.0, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], False),
    "grey_last": ((1, 3, 3, 4, 8), 1.0, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], True),
    "single": ((1, 3, 1, 4, 4), 1.0, [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], False),
}


def make_video(name):
    shape, scale, mean, std, grey = CASES[name]
    v = synth.tensor("video_" + name, shape, scale, 7)
    if grey:
        v[:, :, -1] = 0.0 + 0.01 * v[:, :, -1]          # (0*0.5+0.5)*255 = 127.5 -> in the [117, 137] band
    return v, mean, std


def main():
    captured = {}

    class _Writer:
        def __init__(self, path, **kw):
            self.frames = []
            captured[path] = self

        def append_data(self, frame):
            self.frames.append(np.array(frame))

        def close(self):
            pass

    sys.modules["imageio"] = types.ModuleType("imageio")
    sys.modules["imageio"].get_writer = lambda path, **kw: _Writer(path, **kw)
    sys.path.insert(0, "/root/reference")
    import cv2
    written = {}
    real_imwrite = cv2.imwrite
    cv2.imwrite = lambda path, img, *a: written.__setitem__(path, np.array(img)) or True
    try:
        from utils import video_op
    finally:
        pass
    out = {}
    for name in CASES:
        v, mean, std = make_video(name)
        path = f"/tmp/_golden_{name}.mp4"
        video_op.save_i2vgen_video_safe(path, v.clone(), ["caption"], mean, std, 256)
        if v.shape[2] == 1:
            frames = [written[path + ".png"][:, :, ::-1]]       # the reference writes BGR for cv2
        else:
            frames = captured[path].frames
        out[name + "_frames"] = np.stack(frames) if frames else np.zeros((0,), np.uint8)
        out[name + "_nframes_in"] = np.array(v.shape[2])
    cv2.imwrite = real_imwrite
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "video_out.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
