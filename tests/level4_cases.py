"""Level 4 (FFV1 version 4, micro version 2; -strict experimental) parity cases:
(id, width, height, pix_fmt, encoder options, frames, exact).  `exact`: the reference build's packets are the expected
bytes (RGB layouts).  For planar YUV / gray the reference derives the slice's RCT coefficients -- which only RGB slices
use -- from out-of-bounds reads of the planes as packed RGB (ffv1enc.c:1162-1164, 1103-1112); those cases code the
neutral pair 1, 1 and are checked through the oracle, the reference DECODER and lossless round trips instead."""
LEVEL4_CASES = [
    ("l4_bgr0_range",    96,  80,  "bgr0",        dict(gop=3, level=4, coder=1, slices=4), 4, True),
    ("l4_bgr0_golomb",   96,  80,  "bgr0",        dict(gop=3, level=4, coder=0, slices=4), 4, True),
    ("l4_bgra_range",    101, 67,  "bgra",        dict(gop=2, level=4, coder=1, slices=6), 3, True),
    ("l4_bgra_golomb",   96,  80,  "bgra",        dict(gop=3, level=4, coder=0, slices=4), 3, True),
    ("l4_gbrp9",         96,  80,  "gbrp9le",     dict(gop=3, level=4, slices=4), 3, True),
    ("l4_gbrp14_30sl",   384, 240, "gbrp14le",    dict(gop=4, level=4, coder=2, slices=30), 3, True),
    ("l4_gbrp12_ctx1",   96,  80,  "gbrp12le",    dict(gop=2, level=4, coder=1, context=1, slices=4), 2, True),
    ("l4_yuv420p",       96,  80,  "yuv420p",     dict(gop=3, level=4, coder=1, slices=4), 4, False),
    ("l4_yuv420p_golomb", 96, 80,  "yuv420p",    dict(gop=3, level=4, coder=0, slices=4), 4, False),
    ("l4_yuv444p16",     64,  48,  "yuv444p16le", dict(gop=3, level=4, slices=4), 3, False),
    ("l4_gray",          96,  80,  "gray",        dict(gop=3, level=4, coder=1, slices=4), 3, False),   # plane_count drops to 1 (ffv1enc.c:892)
    ("l4_ya8",           96,  80,  "ya8",         dict(gop=3, level=4, coder=1, slices=4), 3, False),
    ("l4_yuva420p",      96,  80,  "yuva420p",    dict(gop=3, level=4, coder=1, slices=4), 3, False),
]
PCM_FORMATS = ["bgr0", "bgra", "gbrp14le", "yuv420p", "yuv444p16le", "gray", "ya8", "yuva420p"]

def make_frames(case):
    from oracle import synth
    cid, w, h, fmt, opts, n, exact = case
    g = synth.Noisy(w, h, fmt, seed=sum(map(ord, cid)))
    return [g.next() for _ in range(n)]
