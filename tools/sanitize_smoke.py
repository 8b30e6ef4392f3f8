#!/usr/bin/env python3
"""Small encodes / decodes that touch every kernel family, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import numpy as np, ffv1_b200
from oracle import synth

def run(w, h, fmt, n, g, dec=True, up=None, **o):
    gen = synth.Noisy(w, h, fmt, 5)
    frames = [gen.next() for _ in range(n)]
    enc = ffv1_b200.FFV1Encoder(w, h, fmt, g=g, max_batch_frames=n, **o)
    pk = enc.encode_batch(frames)
    if o.get("flags"):
        enc.stats_out()
    if dec:
        d = ffv1_b200.FFV1Decoder(w, h, enc.extradata, max_batch_frames=n)
        out = d.decode_batch([p for p, _ in pk])
        keep = np.ones(len(out[0][0]), bool)
        if fmt == "bgr0":
            keep[3::4] = False
        assert all(np.array_equal(np.asarray(out[i][0])[keep], frames[i].view(np.uint8).reshape(-1)[keep]) for i in range(n))
    print("ok", fmt, o)

run(384, 216, "yuv420p", 6, 4, level=3, coder=1, slices=24)                       # tile-sorted lists, k_pixel_fast, two-warp decode
if os.environ.get("SANITIZE_FIRST_ONLY"):                                         # (racecheck is slow: one case)
    sys.exit(0)
run(704, 96, "yuv420p", 3, 2, level=3, coder=1, slices=4)                         # tall chroma tiles / wide slices
run(192, 108, "yuv422p10le", 4, 4, level=3, coder=0, context=1)                   # large context model lists, five-table decode
run(176, 144, "yuv420p", 4, 3, level=3, coder=0, slices=4)                        # Golomb-Rice lists
run(96, 80, "bgr0", 3, 3, level=4, coder=1, slices=4, strict=-2)                  # level 4: RCT search, header variants
run(96, 80, "gbrp14le", 3, 3, level=4, coder=2, slices=4, strict=-2)
run(176, 144, "yuv420p", 4, 3, level=3, coder=1, slices=4, flags=ffv1_b200.FLAG_PASS1)   # first-pass statistics
rng = np.random.default_rng(1)
for fmt, shapes in (("nv12", [(67, 101), (34, 102)]), ("rgb24", [(67, 303)]), ("yuyv422", [(67, 204)])):
    w, h = 101, 67
    up = ffv1_b200.FFV1Uploader(w, h, fmt, pool_frames=2)
    src = [rng.integers(0, 256, sum(r * b for r, b in shapes), dtype=np.uint8) for _ in range(2)]
    dpl, dls = up.upload(src, shapes)
    enc = ffv1_b200.FFV1Encoder(w, h, up.pix_fmt, g=2, level=3, coder=1, slices=4, max_batch_frames=2)
    ffv1_b200.encode_cuda(enc, dpl, dls, 2)
    print("ok upload", fmt)
