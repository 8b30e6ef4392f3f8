"""Level 4 (version 4, -strict experimental), CPU: the oracle restatement against the reference build and its golden
vectors -- RCT coefficient search, extended slice header, PCM slices (a test knob of the oracle: the reference codes them
only when a slice outgrows its buffer) through the reference DECODER, and the host half of the CUDA codec (extradata)."""
import base64, hashlib, json, os, zlib, numpy as np, pytest
from level4_cases import LEVEL4_CASES, PCM_FORMATS, make_frames
from oracle import ffv1_oracle as O, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "level4.json")))

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

@pytest.mark.parametrize("case", LEVEL4_CASES, ids=[c[0] for c in LEVEL4_CASES])
def test_oracle_against_golden(case):
    cid, w, h, fmt, opts, n, exact = case
    frames = make_frames(case)
    g = GOLD[cid]
    assert md5(b"".join(f.tobytes() for f in frames)) == g["input_md5"]
    o = O.Encoder(w, h, fmt, strict_experimental=1, **opts)
    assert o.extradata.hex() == g["extradata"]
    od = O.Decoder(w, h, fmt, o.extradata)
    for i, f in enumerate(frames):
        pkt, key = o.encode(f)
        if exact:
            assert [len(pkt), md5(pkt), int(key)] == g["packets"][i]
        else:
            assert int(key) == g["packets"][i][2] and abs(len(pkt) - g["packets"][i][0]) <= 4 * 8   # only the coefficient symbols differ
        out = od.decode(pkt)[0]
        keep = np.ones(len(out), bool)
        if fmt == "bgr0":
            keep[3::4] = False
        assert np.array_equal(out[keep], f.view(np.uint8).reshape(-1)[keep])
    if not exact:                                    # the reference's own level-4 stream of this clip decodes as well
        od = O.Decoder(w, h, fmt, bytes.fromhex(g["extradata"]))
        for i, z in enumerate(g["packets_z"]):
            out = od.decode(zlib.decompress(base64.b64decode(z)))[0]
            assert np.array_equal(out, frames[i].view(np.uint8).reshape(-1))

@pytest.mark.parametrize("case", LEVEL4_CASES, ids=[c[0] for c in LEVEL4_CASES])
def test_oracle_against_reference_build(ref, case):
    cid, w, h, fmt, opts, n, exact = case
    frames = make_frames(case)
    r = ref.Encoder(w, h, fmt, strict_experimental=1, **opts)
    o = O.Encoder(w, h, fmt, strict_experimental=1, **opts)
    assert r.extradata == o.extradata
    rd = ref.Decoder(w, h, o.extradata)
    for f in frames:
        a, ka = r.encode(f)
        b, kb = o.encode(f)
        assert ka == kb
        if exact:
            assert a == b
        out, name, k = rd.decode(b)                  # the reference decoder accepts the oracle's stream, losslessly
        keep = np.ones(len(out), bool)
        if fmt == "bgr0":
            keep[3::4] = False
        assert name == fmt and np.array_equal(out[keep], f.view(np.uint8).reshape(-1)[keep])

@pytest.mark.parametrize("fmt", PCM_FORMATS)
def test_pcm_slices_through_the_reference_decoder(ref, fmt):
    w, h = 48, 40
    gen = synth.Noisy(w, h, fmt, 3)
    frames = [gen.next() for _ in range(3)]
    o = O.Encoder(w, h, fmt, gop=2, level=4, coder=1, slices=4, strict_experimental=1, force_pcm=1)
    rd, od = ref.Decoder(w, h, o.extradata), O.Decoder(w, h, fmt, o.extradata)
    for f in frames:
        pkt, key = o.encode(f)
        src = f.view(np.uint8).reshape(-1)
        keep = np.ones(len(src), bool)
        if fmt == "bgr0":
            keep[3::4] = False
        assert np.array_equal(rd.decode(pkt)[0][keep], src[keep])
        assert np.array_equal(od.decode(pkt)[0][keep], src[keep])

def test_level4_needs_strict_experimental():
    import ffv1_b200
    with pytest.raises(ValueError):
        O.Encoder(96, 80, "bgr0", gop=3, level=4, coder=1, slices=4)
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.resolve_encoder(96, 80, "bgr0", g=3, level=4, coder=1, slices=4)
    assert e.value.code == -1094995529               # AVERROR_INVALIDDATA, ffv1enc.c:703-706
    with pytest.raises(ffv1_b200.FFV1Error):         # version 2 stays refused
        ffv1_b200.resolve_encoder(96, 80, "bgr0", g=3, level=2, coder=1, slices=4, strict=-2)

@pytest.mark.parametrize("case", LEVEL4_CASES, ids=[c[0] for c in LEVEL4_CASES])
def test_cuda_codec_extradata(case):
    import ffv1_b200
    cid, w, h, fmt, opts, n, exact = case
    o = dict(opts); o["g"] = o.pop("gop")
    info, xd = ffv1_b200.resolve_encoder(w, h, fmt, strict=-2, **o)
    assert xd.hex() == GOLD[cid]["extradata"] and info.version == 4 and info.micro_version == 2
