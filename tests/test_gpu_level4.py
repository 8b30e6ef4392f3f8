"""Level 4 on the GPU (-m gpu): the CUDA encoder's version-4 streams against the oracle (pinned to the reference build by
tests/test_level4_oracle.py) and the reference build's golden packets; the CUDA decoder on level-4 streams of the CUDA
encoder, of the reference encoder, and on PCM slices."""
import base64, hashlib, json, os, zlib, numpy as np, pytest
from level4_cases import LEVEL4_CASES, PCM_FORMATS, make_frames
from oracle import ffv1_oracle as O, synth

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "level4.json")))

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

def gpu_opts(opts):
    o = dict(opts); o["g"] = o.pop("gop")
    return o

def same(out, f, fmt):
    src = f.view(np.uint8).reshape(-1)
    keep = np.ones(len(src), bool)
    if fmt == "bgr0":
        keep[3::4] = False
    return np.array_equal(np.asarray(out).reshape(-1)[keep], src[keep])

@pytest.mark.parametrize("case", LEVEL4_CASES, ids=[c[0] for c in LEVEL4_CASES])
def test_encode_and_decode(case):
    import ffv1_b200
    cid, w, h, fmt, opts, n, exact = case
    frames = make_frames(case)
    g = GOLD[cid]
    o = O.Encoder(w, h, fmt, strict_experimental=1, **opts)
    enc = ffv1_b200.FFV1Encoder(w, h, fmt, max_batch_frames=2, strict=-2, **gpu_opts(opts))     # batches of 2 split the GOPs
    assert enc.extradata == o.extradata and enc.extradata.hex() == g["extradata"]
    got = enc.encode_batch(frames)
    for i, f in enumerate(frames):
        exp, key = o.encode(f)
        assert got[i][1] == key and got[i][0] == exp, "packet %d differs from the oracle" % i
        if exact:
            assert [len(got[i][0]), md5(got[i][0]), int(key)] == g["packets"][i]                # the reference build's own bytes
    dec = ffv1_b200.FFV1Decoder(w, h, enc.extradata, max_batch_frames=4)
    outs = dec.decode_batch([p for p, _ in got])
    for i, f in enumerate(frames):
        assert same(outs[i][0], f, fmt), "frame %d does not round-trip" % i
    if not exact:                                    # the reference encoder's level-4 stream of the same clip
        dec = ffv1_b200.FFV1Decoder(w, h, bytes.fromhex(g["extradata"]), max_batch_frames=4)
        outs = dec.decode_batch([zlib.decompress(base64.b64decode(z)) for z in g["packets_z"]])
        for i in range(len(outs)):
            assert same(outs[i][0], frames[i], fmt)

@pytest.mark.parametrize("fmt", PCM_FORMATS)
def test_decoder_on_pcm_slices(fmt):
    import ffv1_b200
    w, h = 48, 40
    gen = synth.Noisy(w, h, fmt, 3)
    frames = [gen.next() for _ in range(3)]
    o = O.Encoder(w, h, fmt, gop=2, level=4, coder=1, slices=4, strict_experimental=1, force_pcm=1)
    pkts = [o.encode(f)[0] for f in frames]
    dec = ffv1_b200.FFV1Decoder(w, h, o.extradata, max_batch_frames=4)
    outs = dec.decode_batch(pkts)
    for i, f in enumerate(frames):
        assert same(outs[i][0], f, fmt), "PCM frame %d" % i
        assert outs[i][2] == 0, "PCM frame %d flagged as damaged" % i

def test_mixed_pcm_and_predicted_frames():
    """a PCM slice resets its contexts (ffv1enc.c:1054-1055, the reset flag of ffv1dec.c:346, 419): frames of a GOP that
    alternate between the two modes decode only if the decoder follows the flag"""
    import ffv1_b200
    w, h, fmt = 48, 40, "bgr0"
    gen = synth.Noisy(w, h, fmt, 11)
    frames = [gen.next() for _ in range(4)]
    o = O.Encoder(w, h, fmt, gop=4, level=4, coder=1, slices=4, strict_experimental=1)
    pkts = []
    for i, f in enumerate(frames):
        o.set_force_pcm(i & 1)
        pkts.append(o.encode(f)[0])
    dec = ffv1_b200.FFV1Decoder(w, h, o.extradata, max_batch_frames=4)
    od = O.Decoder(w, h, fmt, o.extradata)
    outs = dec.decode_batch(pkts)
    for i, f in enumerate(frames):
        assert np.array_equal(np.asarray(outs[i][0]).reshape(-1), od.decode(pkts[i])[0]), "frame %d differs from the oracle decoder" % i
        assert same(outs[i][0], f, fmt)
