"""Shared parity-test matrix: (id, width, height, pix_fmt, encoder options, input kind, frames)."""
CASES = [
    # BASELINE.json configs at reduced size (the oracle finishes these in seconds)
    ("c1_cif_intra",        352, 288, "yuv420p",     dict(gop=1,  level=3, coder=1, context=0, slices=4, slicecrc=1), "noisy", 3),
    ("c2_gop_range_24sl",   384, 216, "yuv420p",     dict(gop=4,  level=3, coder=1, context=0, slices=24), "noisy", 6),
    ("c3_422p10_ctx1",      192, 108, "yuv422p10le", dict(gop=4,  level=3, coder=0, context=1), "noisy", 5),
    ("c4_gbrp14_30sl",      384, 240, "gbrp14le",    dict(gop=4,  level=3, coder=2, context=0, slices=30), "noisy", 3),
    # FATE variants (tests/fate/vcodec.mak:113-127)
    ("fate_ffv1_golomb",    352, 288, "yuv420p",     dict(gop=12, slices=4), "noisy", 14),
    ("fate_v0",             352, 288, "yuv420p",     dict(gop=12), "noisy", 14),
    ("fate_v3_444p16",      176, 144, "yuv444p16le", dict(gop=12, level=3), "random", 3),
    ("fate_v3_bgr0",        176, 144, "bgr0",        dict(gop=12, level=3), "noisy", 4),
    # remaining knobs
    ("range_def",           176, 144, "yuv420p",     dict(gop=3, level=3, coder=-2, slices=4), "noisy", 4),
    ("nocrc",               176, 144, "yuv420p",     dict(gop=3, level=3, coder=1, slices=4, slicecrc=0), "noisy", 3),
    ("ctx1_8bit_range",     176, 144, "yuv444p",     dict(gop=3, level=3, coder=1, context=1, slices=4), "noisy", 4),
    ("ctx1_8bit_golomb",    176, 144, "yuv422p",     dict(gop=3, level=3, coder=0, context=1, slices=4), "noisy", 4),
    ("golomb_flat_runs",    352, 288, "yuv420p",     dict(gop=5, level=3, coder=0, slices=6), "flat", 6),
    ("range_flat",          352, 288, "yuv420p",     dict(gop=5, level=3, coder=1, slices=6), "flat", 6),
    ("yuva420p",            176, 144, "yuva420p",    dict(gop=3, level=3, coder=1, slices=4), "noisy", 3),
    ("yuva444p10_golomb",   96,  80,  "yuva444p10le", dict(gop=3, level=3, coder=0, slices=4), "noisy", 3),
    ("gray8",               176, 144, "gray",        dict(gop=3, level=3, coder=0, slices=4), "noisy", 3),
    ("gray16",              96,  80,  "gray16le",    dict(gop=3, level=3, slices=4), "random", 2),
    ("ya8",                 176, 144, "ya8",         dict(gop=3, level=3, coder=1, slices=4), "noisy", 3),
    ("bgra_golomb",         176, 144, "bgra",        dict(gop=3, level=3, coder=0, slices=4), "noisy", 3),
    ("bgra_range_ctx1",     96,  80,  "bgra",        dict(gop=3, level=3, coder=1, context=1, slices=4), "noisy", 3),
    ("gbrp9",               96,  80,  "gbrp9le",     dict(gop=3, level=3, slices=4), "noisy", 3),
    ("gbrp12_random",       96,  80,  "gbrp12le",    dict(gop=2, level=3, slices=4), "random12", 2),
    ("yuv410p_odd",         101, 67,  "yuv410p",     dict(gop=3, level=3, coder=1, slices=6), "noisy", 3),
    ("yuv411p_odd_golomb",  101, 67,  "yuv411p",     dict(gop=3, level=3, coder=0, slices=4), "noisy", 3),
    ("yuv440p",             100, 66,  "yuv440p",     dict(gop=3, level=3, coder=1, slices=4), "noisy", 3),
    ("vsynth3_size",        34,  34,  "yuv420p",     dict(gop=12, slices=4), "random", 5),
    ("v1_10bit",            96,  80,  "yuv420p10le", dict(gop=3, level=1), "noisy", 4),
    ("v0_range",            96,  80,  "yuv420p",     dict(gop=3, level=0, coder=1), "noisy", 4),
    ("v1_golomb_ctx1",      96,  80,  "yuv444p",     dict(gop=3, level=1, coder=0, context=1), "noisy", 4),
    ("slices9_random8",     99,  99,  "yuv444p",     dict(gop=2, level=3, coder=1, slices=9), "random", 2),
    ("default_big_autov3",  736, 580, "yuv420p",     dict(gop=2), "flat", 2),
    ("tiny_1slice",         16,  8,   "yuv420p",     dict(gop=2, level=3, coder=1, slices=1), "random", 2),
]

def make_frames(case):
    from oracle import synth
    cid, w, h, fmt, opts, kind, n = case
    if kind == "noisy":
        g = synth.Noisy(w, h, fmt, seed=abs(hash(cid)) % 1000 if False else sum(map(ord, cid)))
        return [g.next() for _ in range(n)]
    if kind == "flat":
        return [synth.flat_bars(w, h, fmt, i) for i in range(n)]
    if kind == "random":
        return [synth.random_frame(w, h, fmt, 100 + i) for i in range(n)]
    if kind == "random12":
        return [synth.random_frame(w, h, fmt, 100 + i, maxval=4095) for i in range(n)]
    raise ValueError(kind)
