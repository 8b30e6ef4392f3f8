"""The reference's own FATE goldens for this path (tests/ref/vsynth/vsynth{1,2,3}-ffv1{,-v0,-v3-yuv420p},
tests/fate/vcodec.mak:113-118): AVI MD5 + size reproduced through oracle/_ref (reference encoder + the reference's
AVI muxer), and the oracle restatement matching the reference packet by packet on the very same 50-frame clips.
Needs /root/reference (present in the build container only) -> skipped on the GPU box.
The nine 10/16-bit/bgr0 goldens (ffv1-v3-yuv422p10 / -yuv444p16 / -bgr0) take the clip through the reference's own
libswscale first, the way ffmpeg's auto-inserted scale filter does (oracle/ref_harness.c:ffv1ref_sws_convert): all 18
goldens of this path that can be reproduced without external samples are reproduced."""
import hashlib, os, numpy as np, pytest
from oracle import ffv1_oracle as O

REF_TREE = "/root/reference"
VARIANTS = [("ffv1", dict(slices=4), -1, 4), ("ffv1-v0", dict(), -1, 0), ("ffv1-v3-yuv420p", dict(level=3), 3, 0)]

@pytest.mark.skipif(not os.path.isdir(REF_TREE), reason="reference tree not mounted")
@pytest.mark.parametrize("clip", ["vsynth1", "vsynth2", "vsynth3"])
def test_fate_vsynth(ref, clip):
    raw, w, h = ref.vsynth(clip)
    fsz = w * h * 3 // 2
    n = len(raw) // fsz
    assert n == 50
    for name, opts, level, slices in VARIANTS:
        gold = open(os.path.join(REF_TREE, "tests/ref/vsynth/%s-%s" % (clip, name))).read().split("\n")
        avi = ref.fate_avi(raw, n, w, h, "yuv420p", level, slices)
        assert hashlib.md5(avi).hexdigest() == gold[0].split()[0], "AVI md5 differs from FATE golden"
        assert len(avi) == int(gold[1].split()[0])
        assert hashlib.md5(raw.tobytes()).hexdigest() == gold[2].split()[0]      # lossless: decoded md5 == source md5
        # oracle == reference on all 50 packets (gop 12 default -> keyframes 0,12,24,36,48 + carried-over state)
        r, o = ref.Encoder(w, h, "yuv420p", **opts), O.Encoder(w, h, "yuv420p", **opts)
        od = O.Decoder(w, h, "yuv420p", o.extradata)
        assert r.extradata == o.extradata
        for i in range(n):
            f = raw[i * fsz:(i + 1) * fsz]
            a, ka = r.encode(f)
            b, kb = o.encode(f)
            assert a == b and ka == kb == (i % 12 == 0), (name, i)
            assert a in avi                                   # the packet is what the muxer stored
            out, _, dmg = od.decode(a)
            assert np.array_equal(out, f) and dmg == 0

CONVERTED = [("ffv1-v3-yuv422p10", "yuv422p10le"), ("ffv1-v3-yuv444p16", "yuv444p16le"), ("ffv1-v3-bgr0", "bgr0")]

@pytest.mark.skipif(not os.path.isdir(REF_TREE), reason="reference tree not mounted")
@pytest.mark.parametrize("clip", ["vsynth1", "vsynth2", "vsynth3"])
def test_fate_vsynth_converted_formats(ref, clip):
    """tests/fate/vcodec.mak:119-127: -level 3 -pix_fmt yuv422p10 / yuv444p16 / bgr0 (range coder forced above 8 bits,
    Golomb-Rice + RCT for bgr0); oracle == reference on the first 14 packets (keyframe 0, 12 + carried-over state)"""
    raw, w, h = ref.vsynth(clip)
    for name, fmt in CONVERTED:
        conv = ref.sws_convert(raw, w, h, fmt)
        fsz = len(conv) // 50
        gold = open(os.path.join(REF_TREE, "tests/ref/vsynth/%s-%s" % (clip, name))).read().split("\n")
        avi = ref.fate_avi(conv, 50, w, h, fmt, 3, 0)
        assert hashlib.md5(avi).hexdigest() == gold[0].split()[0], "AVI md5 differs from FATE golden"
        assert len(avi) == int(gold[1].split()[0])
        r, o = ref.Encoder(w, h, fmt, level=3), O.Encoder(w, h, fmt, level=3)
        od = O.Decoder(w, h, fmt, o.extradata)
        assert r.extradata == o.extradata
        for i in range(14):
            f = conv[i * fsz:(i + 1) * fsz]
            a, ka = r.encode(f)
            b, kb = o.encode(f)
            assert a == b and ka == kb == (i % 12 == 0), (name, i)
            out, _, dmg = od.decode(a)
            if fmt == "bgr0":                                 # the unused byte is not coded: swscale writes 255 there, a decoder 0 (ffv1dec.c:270-276)
                out, f = out.reshape(-1, 4)[:, :3], f.reshape(-1, 4)[:, :3]
            assert np.array_equal(out, f) and dmg == 0

@pytest.mark.skipif(not os.path.isdir(REF_TREE), reason="reference tree not mounted")
def test_nut_round_trip_reference_pair(ref):
    """the container path of SURVEY 8(f) rank 2 with the reference's own codec: ffv1 -> nutenc.c -> nutdec.c -> ffv1"""
    raw, w, h = ref.vsynth("vsynth1")
    nut = ref.mux("nut", "ffv1", raw, 50, w, h, "yuv420p", level=3, slices=4)
    assert bytes(nut[:24]) == b"nut/multimedia container"
    out, n, fmt = ref.nut_decode("ffv1", nut, len(raw))
    assert n == 50 and fmt == "yuv420p" and np.array_equal(out, raw)
