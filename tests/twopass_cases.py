"""Two-pass parity cases (AV_CODEC_FLAG_PASS1 / _PASS2, ffv1enc.c:898-986, 1235-1277):
(id, width, height, pix_fmt, encoder options, frames)."""
TWOPASS_CASES = [
    ("tp_420_custom",   176, 144, "yuv420p",     dict(gop=3, level=3, coder=1,  context=0, slices=4), 6),
    ("tp_420_default",  176, 144, "yuv420p",     dict(gop=3, level=3, coder=-2, context=0, slices=4), 6),
    ("tp_444_ctx1",     96,  80,  "yuv444p",     dict(gop=2, level=3, coder=1,  context=1, slices=4), 4),
    ("tp_422p10",       96,  80,  "yuv422p10le", dict(gop=3, coder=0, context=0, slices=4), 4),      # range coder forced, level unset -> 3
    ("tp_bgr0",         96,  80,  "bgr0",        dict(gop=3, level=3, coder=1, slices=4), 4),
    ("tp_golomb",       96,  80,  "yuv420p",     dict(gop=3, level=3, coder=0, slices=4), 4),         # no statistics in Golomb-Rice mode
]

def make_frames(case):
    from oracle import synth
    cid, w, h, fmt, opts, n = case
    g = synth.Noisy(w, h, fmt, seed=sum(map(ord, cid)))
    return [g.next() for _ in range(n)]
