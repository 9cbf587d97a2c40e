"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's video write-out arithmetic
(utils/video_op.py:167-213 save_i2vgen_video_safe; the same lines in save_t2vhigen_video_safe :263-309).

Pinned against the REAL reference function: oracle/make_golden_video.py runs save_i2vgen_video_safe itself (imageio
replaced by a frame-capturing writer, since imageio/ffmpeg are absent) and freezes the frames it hands to the encoder
into tests/golden/video_out.npz; tests/test_oracle_golden.py checks this file against those frames bit for bit.
"""
from __future__ import annotations

import numpy as np


def frames_uint8(gen_video, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """gen_video float32 [b, 3, f, h, w] -> list of f uint8 [h, w, 3] frames of batch entry 0 (:180-192):
    mul_(std), add_(mean) (two fp32 roundings), clamp_(0, 1), * 255.0, 'b c f h w -> b f h w c', astype('uint8')."""
    v = np.asarray(gen_video, dtype=np.float32)
    m = np.asarray(mean, dtype=np.float32).reshape(1, -1, 1, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, -1, 1, 1, 1)
    v = (v * s).astype(np.float32)
    v = (v + m).astype(np.float32)
    v = np.clip(v, np.float32(0), np.float32(1))
    v = (v * np.float32(255.0)).astype(np.float32)
    imgs = np.transpose(v, (0, 2, 3, 4, 1))[0]
    return [img.astype("uint8") for img in imgs]


def drop_anomalous_last_frame(frames):
    """:196-202: the last frame is skipped when more than 40 % of its bytes lie in [117, 137]."""
    if len(frames) <= 1:
        return list(frames)
    last = frames[-1]
    ratio = np.sum((last >= 117) & (last <= 137)) / last.size
    return list(frames[:-1]) if ratio > 0.4 else list(frames)


def band_counts(frames):
    return [int(np.sum((f >= 117) & (f <= 137))) for f in frames]
