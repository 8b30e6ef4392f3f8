"""Drop-in test at the reference's plugin boundary (run with -m gpu): the AVCodec shim integration/ffv1_b200_avcodec.c is
registered with the REFERENCE's own libavcodec (oracle/_ref/libffv1ref.so, unmodified sources) via avcodec_register and
driven through avcodec_open2 / avcodec_encode_video2 / avcodec_decode_video2 with the same AVOptions as "ffv1".
Packets and extradata must be byte-identical to what the reference encoder produces in the same process, and the
reference decoder / our decoder must both give back the source frames."""
import ctypes, os, numpy as np, pytest
from cases import CASES, make_frames
from oracle import ffv1_ref, pixfmt
from oracle.ffv1_oracle import split_planes, _plane_args

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "integration", "_build", "libffv1_b200_avcodec.so")

@pytest.fixture(scope="module")
def lavc():
    if not ffv1_ref.available() or not os.path.exists(SHIM):
        pytest.skip("reference build / shim not present (built only where /root/reference exists)")
    import ffv1_b200
    assert ffv1_b200.device_count() >= 1
    ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libffv1ref.so"), mode=ctypes.RTLD_GLOBAL)   # the reference's lavc/lavu symbols
    L = ffv1_ref.lib()
    shim = ctypes.CDLL(SHIM, mode=ctypes.RTLD_GLOBAL)
    L.ffv1ref_register_codec.argtypes = [ctypes.c_void_p]
    L.ffv1ref_register_codec(ctypes.addressof(ctypes.c_char.in_dll(shim, "ff_ffv1_b200_encoder")))
    L.ffv1ref_register_codec(ctypes.addressof(ctypes.c_char.in_dll(shim, "ff_ffv1_b200_decoder")))
    L.ffv1ref_enc_open_named.restype = ctypes.c_void_p
    L.ffv1ref_enc_open_named.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 7
    L.ffv1ref_dec_open_named.restype = ctypes.c_void_p
    L.ffv1ref_dec_open_named.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return L

def encode_named(L, name, case, batch, props=None):
    """props: optional function frame index -> (sar_num, sar_den, interlaced, top_field_first)"""
    cid, w, h, fmt, opts, kind, n = case
    o = dict(gop=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1); o.update(opts)
    hnd = L.ffv1ref_enc_open_named(name.encode(), w, h, fmt.encode(), o["gop"], o["level"], o["coder"], o["context"],
                                   o["slices"], o["slicecrc"], batch)
    assert hnd, "avcodec_open2(%s) failed" % name
    ed = ctypes.create_string_buffer(1 << 16)
    ned = L.ffv1ref_enc_extradata(hnd, ed, 1 << 16)
    cap = 65536 + pixfmt.frame_bytes(fmt, w, h) * 4
    buf = ctypes.create_string_buffer(cap)
    key = ctypes.c_int()
    pkts = []
    for fi, f in enumerate(make_frames(case)):
        planes = split_planes(np.ascontiguousarray(f).view(np.uint8).reshape(-1), fmt, w, h)
        ptrs, strides = _plane_args(planes)
        sn, sd, il, tff = props(fi) if props else (0, 1, 0, 0)
        r = L.ffv1ref_enc_frame(hnd, ptrs, strides, sn, sd, il, tff, buf, cap, ctypes.byref(key))
        assert r >= 0
        if r:
            pkts.append((buf.raw[:r], bool(key.value)))
    while True:                                            # drain (frame = NULL), as ffmpeg.c flush_encoders does
        r = L.ffv1ref_enc_frame(hnd, None, None, 0, 1, 0, 0, buf, cap, ctypes.byref(key))
        assert r >= 0
        if not r:
            break
        pkts.append((buf.raw[:r], bool(key.value)))
    L.ffv1ref_enc_close(hnd)
    return ed.raw[:max(ned, 0)], pkts

DROPIN = [c for c in CASES if c[0] in ("c1_cif_intra", "c2_gop_range_24sl", "c3_422p10_ctx1", "c4_gbrp14_30sl",
                                       "fate_ffv1_golomb", "fate_v3_bgr0", "yuva420p", "range_def")]

@pytest.mark.parametrize("case", DROPIN, ids=[c[0] for c in DROPIN])
def test_same_packets_through_avcodec_api(lavc, case):
    ref_ed, ref_pkts = encode_named(lavc, "ffv1", case, 0)
    our_ed, our_pkts = encode_named(lavc, "ffv1_b200", case, 5)           # batches of 5 frames: delayed output + drain
    assert our_ed == ref_ed
    assert len(our_pkts) == len(ref_pkts) == case[6]
    for i, (a, b) in enumerate(zip(our_pkts, ref_pkts)):
        assert a[1] == b[1], "AV_PKT_FLAG_KEY of packet %d" % i
        assert a[0] == b[0], "packet %d differs from the reference encoder's" % i

PROPS = [(1, 1, 0, 0), (1, 1, 0, 0), (4, 3, 1, 1), (4, 3, 1, 1), (4, 3, 1, 0), (0, 1, 0, 0), (16, 11, 0, 0), (16, 11, 1, 1)]

@pytest.mark.parametrize("cid,batch", [("fate_ffv1_golomb", 5), ("fate_ffv1_golomb", 3), ("c2_gop_range_24sl", 4), ("c2_gop_range_24sl", 2),
                                       ("c4_gbrp14_30sl", 2)])
def test_sar_and_field_order_change_mid_stream(lavc, cid, batch):
    """AVFrame.sample_aspect_ratio / interlaced_frame / top_field_first change while frames are queued and while packets of
    earlier batches are still being handed out (ffv1enc.c:1044-1049 codes them per frame): same packets as the reference"""
    case = [c for c in CASES if c[0] == cid][0]
    pr = lambda i: PROPS[i % len(PROPS)]
    ref_ed, ref_pkts = encode_named(lavc, "ffv1", case, 0, pr)
    our_ed, our_pkts = encode_named(lavc, "ffv1_b200", case, batch, pr)
    plain = encode_named(lavc, "ffv1", case, 0)[1]
    assert plain[0][0] != ref_pkts[0][0], "the properties must change the bytes"
    assert our_ed == ref_ed and len(our_pkts) == len(ref_pkts) == case[6]
    for i, (a, b) in enumerate(zip(our_pkts, ref_pkts)):
        assert a == b, "packet %d differs from the reference encoder's" % i

@pytest.mark.parametrize("batch", [1, 4, 64])
def test_batched_decode_through_avcodec_api(lavc, batch):
    """the decoder's "batch" option: pictures come out batch-1 calls late and the rest on the drain (AV_CODEC_CAP_DELAY)"""
    case = [c for c in CASES if c[0] == "fate_ffv1_golomb"][0]
    cid, w, h, fmt, opts, kind, n = case
    ed, pkts = encode_named(lavc, "ffv1", case, 0)
    lavc.ffv1ref_dec_open_named_opts.restype = ctypes.c_void_p
    lavc.ffv1ref_dec_open_named_opts.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
    hnd = lavc.ffv1ref_dec_open_named_opts(b"ffv1_b200", w, h, ed, len(ed), ("batch=%d" % batch).encode())
    assert hnd
    cap = w * h * 8 + 64
    out = np.zeros(cap, np.uint8); name = ctypes.create_string_buffer(32); key = ctypes.c_int()
    frames = [np.ascontiguousarray(f).view(np.uint8).reshape(-1) for f in make_frames(case)]
    got = []
    for i in range(n):
        r = lavc.ffv1ref_dec_packet(hnd, pkts[i][0], len(pkts[i][0]), out.ctypes.data, cap, name, ctypes.byref(key))
        assert r >= 0
        if r:
            got.append((out[:r].copy(), bool(key.value)))
    assert len(got) == (n - (batch - 1) if n >= batch else 0)
    while True:
        r = lavc.ffv1ref_dec_packet(hnd, b"", 0, out.ctypes.data, cap, name, ctypes.byref(key))
        assert r >= 0
        if not r:
            break
        got.append((out[:r].copy(), bool(key.value)))
    lavc.ffv1ref_dec_close(hnd)
    assert len(got) == n
    for i in range(n):
        assert got[i][1] == pkts[i][1] and np.array_equal(got[i][0], frames[i]), "picture %d" % i

def test_bench_harness_loop_gives_the_same_packets(lavc):
    """tools/bench_avcodec.py's loop (zero-copy AVFrames, worker-thread staging, pipelined batches): packets of -c:v ffv1_b200
    == packets of -c:v ffv1 on a tiled clip, including batches that end inside a GOP"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_avcodec as A
    case = [c for c in CASES if c[0] == "c2_gop_range_24sl"][0]
    cid, w, h, fmt, opts, kind, n = case
    clip = np.stack([np.ascontiguousarray(f).view(np.uint8).reshape(-1) for f in make_frames(case)])
    o = dict(opts); gop = o.pop("gop")
    _, nb_ref, ref = A.encode("ffv1", clip, w, h, fmt, 50, gop, o, nkeep=50)
    for extra in ("batch=7:copy_threads=3", "batch=16:copy_threads=0", "batch=64:copy_threads=8"):
        _, nb, got = A.encode("ffv1_b200", clip, w, h, fmt, 50, gop, o, extra=extra, nkeep=50)
        assert nb == nb_ref and got == ref, extra

def test_decode_through_avcodec_api(lavc):
    case = [c for c in CASES if c[0] == "c2_gop_range_24sl"][0]
    cid, w, h, fmt, opts, kind, n = case
    ed, pkts = encode_named(lavc, "ffv1", case, 0)
    hnd = lavc.ffv1ref_dec_open_named(b"ffv1_b200", w, h, ed, len(ed))
    assert hnd
    cap = w * h * 8 + 64
    out = np.zeros(cap, np.uint8); name = ctypes.create_string_buffer(32); key = ctypes.c_int()
    for i, f in enumerate(make_frames(case)):
        r = lavc.ffv1ref_dec_packet(hnd, pkts[i][0], len(pkts[i][0]), out.ctypes.data, cap, name, ctypes.byref(key))
        assert r == pixfmt.frame_bytes(fmt, w, h) and name.value.decode() == fmt and bool(key.value) == pkts[i][1]
        assert np.array_equal(out[:r], np.ascontiguousarray(f).view(np.uint8).reshape(-1))
    lavc.ffv1ref_dec_close(hnd)

@pytest.mark.parametrize("cid", ["c2_gop_range_24sl", "fate_ffv1_golomb", "c3_422p10_ctx1"])
def test_cuda_frames_through_avcodec_api(lavc, cid):
    """avctx->pix_fmt = AV_PIX_FMT_CUDA: AVFrames whose data[] are device pointers (sw_format from hw_frames_ctx) must give
    the packets the reference encoder produces from the same pixels in system memory"""
    import torch
    case = [c for c in CASES if c[0] == cid][0]
    _, w, h, fmt, opts, kind, n = case
    ref_ed, ref_pkts = encode_named(lavc, "ffv1", case, 0)
    o = dict(gop=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1); o.update(opts)
    lavc.ffv1ref_enc_open_named_cuda.restype = ctypes.c_void_p
    lavc.ffv1ref_enc_open_named_cuda.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 7
    lavc.ffv1ref_enc_frame_cuda.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                            ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    hnd = lavc.ffv1ref_enc_open_named_cuda(b"ffv1_b200", w, h, fmt.encode(), o["gop"], o["level"], o["coder"], o["context"],
                                           o["slices"], o["slicecrc"], 4)
    assert hnd, "avcodec_open2(ffv1_b200, AV_PIX_FMT_CUDA) failed"
    ed = ctypes.create_string_buffer(1 << 16)
    ned = lavc.ffv1ref_enc_extradata(hnd, ed, 1 << 16)
    assert ed.raw[:ned] == ref_ed
    cap = 65536 + pixfmt.frame_bytes(fmt, w, h) * 4
    buf = ctypes.create_string_buffer(cap); key = ctypes.c_int()
    keep, pkts = [], []
    for f in make_frames(case):
        planes = split_planes(np.ascontiguousarray(f).view(np.uint8).reshape(-1), fmt, w, h)
        ptrs, strides = (ctypes.c_void_p * 4)(), (ctypes.c_int * 4)()
        for i, pl in enumerate(planes):
            rows, rb = pl.shape[0], pl.shape[1] * pl.itemsize if pl.ndim == 2 else pl.shape[1] * pl.shape[2] * pl.itemsize
            pitch = (rb + 255) & ~255                                       # device frames with a pitch, like cuMemAllocPitch
            d = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda")
            d[:, :rb] = torch.from_numpy(np.ascontiguousarray(pl).view(np.uint8).reshape(rows, rb)).cuda()
            keep.append(d)
            ptrs[i] = d.data_ptr(); strides[i] = pitch
        torch.cuda.synchronize()
        r = lavc.ffv1ref_enc_frame_cuda(hnd, ptrs, strides, buf, cap, ctypes.byref(key))
        assert r >= 0
        if r:
            pkts.append((buf.raw[:r], bool(key.value)))
    while True:
        r = lavc.ffv1ref_enc_frame(hnd, None, None, 0, 1, 0, 0, buf, cap, ctypes.byref(key))
        assert r >= 0
        if not r:
            break
        pkts.append((buf.raw[:r], bool(key.value)))
    lavc.ffv1ref_enc_close(hnd)
    assert len(pkts) == len(ref_pkts)
    for i, (a, b) in enumerate(zip(pkts, ref_pkts)):
        assert a == b, "packet %d differs from the reference encoder's" % i

FATE = [("ffv1", -1, 4, "yuv420p"), ("ffv1-v0", -1, 0, "yuv420p"), ("ffv1-v3-yuv420p", 3, 0, "yuv420p"),      # tests/fate/vcodec.mak:113-127
        ("ffv1-v3-yuv422p10", 3, 0, "yuv422p10le"), ("ffv1-v3-yuv444p16", 3, 0, "yuv444p16le"), ("ffv1-v3-bgr0", 3, 0, "bgr0")]

@pytest.mark.parametrize("clip", ["vsynth1", "vsynth2", "vsynth3"])
def test_fate_goldens_through_the_dropin(lavc, clip):
    """FATE's enc_dec procedure (tests/fate-run.sh:171-193) with -c:v ffv1_b200: the reference's synthetic clips (videogen /
    rotozoom, converted by the reference's libswscale for the 10/16-bit/bgr0 variants), our encoder behind the reference's
    libavcodec, the reference's AVI muxer -> the file must hash to the reference's own golden: all 18 goldens of the path
    that need no external sample."""
    import hashlib, json
    try:
        raw, w, h = ffv1_ref.vsynth(clip)
    except Exception as ex:
        pytest.skip("clip generator not available here: %r" % ex)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fate_vsynth.json")))
    n = len(raw) // (w * h * 3 // 2)
    assert n == 50
    lavc.ffv1ref_fate_avi_named.restype = ctypes.c_int64
    lavc.ffv1ref_fate_avi_named.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    for name, level, slices, fmt in FATE:
        src = raw if fmt == "yuv420p" else ffv1_ref.sws_convert(raw, w, h, fmt)
        out = np.zeros(len(src) + (4 << 20), np.uint8)
        size = lavc.ffv1ref_fate_avi_named(b"ffv1_b200", 16, src.ctypes.data, n, w, h, fmt.encode(), level, slices, out.ctypes.data, len(out))
        g = gold["%s-%s" % (clip, name)]
        assert size == g["avi_size"], (clip, name, size)
        assert hashlib.md5(out[:size].tobytes()).hexdigest() == g["avi_md5"], (clip, name)
        if fmt == "yuv420p":
            assert hashlib.md5(raw.tobytes()).hexdigest() == g["decoded_md5"]

@pytest.mark.parametrize("fmt,level,slices", [("yuv420p", 3, 4), ("yuv420p", -1, 0), ("yuv422p10le", 3, 0), ("bgr0", 3, 4)])
def test_nut_round_trip_through_the_dropin(lavc, fmt, level, slices):
    """container round trip: -c:v ffv1_b200 -> the reference's NUT muxer -> the reference's NUT
    demuxer -> ffv1_b200 decoder (one picture per packet, and batched with delayed output) and the reference decoder:
    the file is byte-identical to the one the reference encoder produces, and every decoder gives the source back"""
    try:
        raw, w, h = ffv1_ref.vsynth("vsynth1")
    except Exception as ex:
        pytest.skip("clip generator not available here: %r" % ex)
    src = raw if fmt == "yuv420p" else ffv1_ref.sws_convert(raw, w, h, fmt)
    ours = ffv1_ref.mux("nut", "ffv1_b200", src, 50, w, h, fmt, level=level, slices=slices, batch=16)
    theirs = ffv1_ref.mux("nut", "ffv1", src, 50, w, h, fmt, level=level, slices=slices)
    assert np.array_equal(ours, theirs), "NUT file differs from the one written with the reference encoder"
    want = src.reshape(50, -1)
    if fmt == "bgr0":
        want = want.reshape(50, -1, 4).copy(); want[:, :, 3] = 0          # the unused byte is not coded (ffv1dec.c:270-276)
        want = want.reshape(50, -1)
    for dec, opts in (("ffv1_b200", ""), ("ffv1_b200", "batch=16"), ("ffv1", "")):
        out, n, name = ffv1_ref.nut_decode(dec, ours, len(src), opts)
        assert n == 50 and name == fmt, (dec, opts, n, name)
        assert np.array_equal(out.reshape(50, -1), want), (dec, opts)

def _run_named(L, name, w, h, fmt, opts, frames, batch, strict=0, two_pass=0, stats_in=None, want_stats=False):
    L.ffv1ref_enc_open_named_ex.restype = ctypes.c_void_p
    L.ffv1ref_enc_open_named_ex.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 9 + [ctypes.c_char_p]
    L.ffv1ref_enc_stats_out.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    o = dict(gop=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1); o.update(opts)
    if isinstance(stats_in, str):
        stats_in = stats_in.encode()
    hnd = L.ffv1ref_enc_open_named_ex(name.encode(), w, h, fmt.encode(), o["gop"], o["level"], o["coder"], o["context"],
                                      o["slices"], o["slicecrc"], batch, strict, two_pass, stats_in)
    assert hnd, "avcodec_open2(%s) failed" % name
    ed = ctypes.create_string_buffer(1 << 16)
    ned = L.ffv1ref_enc_extradata(hnd, ed, 1 << 16)
    cap = 65536 + pixfmt.frame_bytes(fmt, w, h) * 4
    buf = ctypes.create_string_buffer(cap)
    key = ctypes.c_int()
    pkts = []
    for f in frames:
        planes = split_planes(np.ascontiguousarray(f).view(np.uint8).reshape(-1), fmt, w, h)
        ptrs, strides = _plane_args(planes)
        r = L.ffv1ref_enc_frame(hnd, ptrs, strides, 0, 1, 0, 0, buf, cap, ctypes.byref(key))
        assert r >= 0
        if r:
            pkts.append(buf.raw[:r])
    while True:
        r = L.ffv1ref_enc_frame(hnd, None, None, 0, 1, 0, 0, buf, cap, ctypes.byref(key))
        assert r >= 0
        if not r:
            break
        pkts.append(buf.raw[:r])
    stats = None
    if want_stats:
        sb = ctypes.create_string_buffer(8 << 20)
        n = L.ffv1ref_enc_stats_out(hnd, sb, 8 << 20)
        assert n > 0
        stats = sb.value.decode()
    L.ffv1ref_enc_close(hnd)
    return ed.raw[:max(ned, 0)], pkts, stats

def test_two_pass_through_avcodec_api(lavc):
    """-pass 1 / -pass 2 with -c:v ffv1_b200: stats_out after the flush and the second pass's stream equal the reference
    codec's (AV_CODEC_FLAG_PASS1/2 and stats_in reach the library through the shim)"""
    from twopass_cases import TWOPASS_CASES, make_frames as tp_frames
    case = TWOPASS_CASES[0]
    cid, w, h, fmt, opts, n = case
    frames = tp_frames(case)
    ed_r, p_r, st_r = _run_named(lavc, "ffv1", w, h, fmt, opts, frames, 0, two_pass=1, want_stats=True)
    ed_g, p_g, st_g = _run_named(lavc, "ffv1_b200", w, h, fmt, opts, frames, 4, two_pass=1, want_stats=True)
    assert ed_r == ed_g and p_r == p_g and st_r == st_g
    ed_r2, p_r2, _ = _run_named(lavc, "ffv1", w, h, fmt, opts, frames, 0, two_pass=2, stats_in=st_r)
    ed_g2, p_g2, _ = _run_named(lavc, "ffv1_b200", w, h, fmt, opts, frames, 4, two_pass=2, stats_in=st_g)
    assert ed_r2 == ed_g2 and p_r2 == p_g2 and len(ed_r2) > len(ed_r)

def test_level4_through_avcodec_api(lavc):
    """-level 4 -strict experimental with -c:v ffv1_b200 on RGB content: the reference codec's own bytes; without -strict
    both refuse"""
    from level4_cases import LEVEL4_CASES, make_frames as l4_frames
    for case in [c for c in LEVEL4_CASES if c[0] in ("l4_bgr0_range", "l4_bgra_golomb", "l4_gbrp9")]:
        cid, w, h, fmt, opts, n, exact = case
        frames = l4_frames(case)
        ed_r, p_r, _ = _run_named(lavc, "ffv1", w, h, fmt, opts, frames, 0, strict=1)
        ed_g, p_g, _ = _run_named(lavc, "ffv1_b200", w, h, fmt, opts, frames, 2, strict=1)
        assert ed_r == ed_g and p_r == p_g, cid
    lavc.ffv1ref_enc_open_named_ex.restype = ctypes.c_void_p
    for name in (b"ffv1", b"ffv1_b200"):
        assert not lavc.ffv1ref_enc_open_named_ex(name, 96, 80, b"bgr0", 3, 4, 1, 0, 4, -1, 2, 0, 0, None)

@pytest.mark.parametrize("fmt,level,slices", [("yuv420p", 3, 4), ("yuv422p10le", 3, 0), ("bgr0", 3, 4)])
def test_matroska_file_through_the_dropin(lavc, fmt, level, slices):
    """SURVEY 8(f) rank 2, Matroska (libavformat/matroskaenc.c, codec private = BITMAPINFOHEADER + extradata, riff.c:316):
    the file written from -c:v ffv1_b200 is byte-identical to the one written from the reference encoder"""
    try:
        raw, w, h = ffv1_ref.vsynth("vsynth1")
    except Exception as ex:
        pytest.skip("clip generator not available here: %r" % ex)
    src = raw if fmt == "yuv420p" else ffv1_ref.sws_convert(raw, w, h, fmt)
    ours = ffv1_ref.mux("matroska", "ffv1_b200", src, 50, w, h, fmt, level=level, slices=slices, batch=16)
    theirs = ffv1_ref.mux("matroska", "ffv1", src, 50, w, h, fmt, level=level, slices=slices)
    assert bytes(ours[:4]) == b"\x1a\x45\xdf\xa3" and len(ours) > 50 * 1000
    assert np.array_equal(ours, theirs), "Matroska file differs from the one written with the reference encoder"
