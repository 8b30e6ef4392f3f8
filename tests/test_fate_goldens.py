"""The reference's own FATE goldens for this path (tests/ref/vsynth/vsynth{1,2,3}-ffv1{,-v0,-v3-yuv420p},
tests/fate/vcodec.mak:113-118): AVI MD5 + size reproduced through oracle/_ref (reference encoder + the reference's
AVI muxer), and the oracle restatement matching the reference packet by packet on the very same 50-frame clips.
Needs /root/reference (present in the build container only) -> skipped on the GPU box.
The six 10/16-bit/bgr0 goldens need libswscale to convert the input and are not reproduced here; those
formats are pinned by test_oracle_vs_ref.py instead."""
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
