"""GPU parity tests of the decoder (run with -m gpu on the B200 box): packets produced by the oracle encoder (pinned to
the reference build) are decoded by the CUDA decoder through the C ABI; frames must equal the oracle decoder's output
byte for byte (and the source frames wherever the slices cover them), for range-coder and Golomb-Rice streams,
state-carry-over non-keyframes, batches that split GOPs, and damaged slices."""
import numpy as np, pytest
from cases import CASES, make_frames
from oracle import ffv1_oracle as O

pytestmark = pytest.mark.gpu
V3 = CASES          # every stream version: 0/1 (parameters in-band, no extradata) and 3

@pytest.fixture(scope="module")
def B():
    import ffv1_b200
    assert ffv1_b200.device_count() >= 1
    return ffv1_b200

def oracle_stream(case):
    cid, w, h, fmt, opts, kind, n = case
    frames = make_frames(case)
    enc = O.Encoder(w, h, fmt, **opts)
    pkts = [enc.encode(f)[0] for f in frames]
    dec = O.Decoder(w, h, fmt, enc.extradata)
    ref = [dec.decode(p) for p in pkts]
    return frames, enc.extradata, pkts, ref

@pytest.mark.parametrize("case", V3, ids=[c[0] for c in V3])
def test_decode_matches_oracle(B, case):
    cid, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    d = B.FFV1Decoder(w, h, extradata, max_batch_frames=4)                 # batches of 4 split the GOPs
    got = d.decode_batch(pkts)
    assert d.pix_fmt == fmt
    for i in range(len(pkts)):
        assert got[i][1] == ref[i][1], "key flag of frame %d" % i
        assert got[i][2] == 0 and ref[i][2] == 0
        assert np.array_equal(got[i][0], ref[i][0]), "frame %d differs from the oracle decoder (first byte %d)" % (
            i, int(np.argmax(got[i][0] != ref[i][0])))

def test_roundtrip_is_lossless(B):
    case = [c for c in CASES if c[0] == "c2_gop_range_24sl"][0]
    cid, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    got = B.FFV1Decoder(w, h, extradata, max_batch_frames=8).decode_batch(pkts)
    for i, f in enumerate(frames):
        assert np.array_equal(got[i][0], np.ascontiguousarray(f).view(np.uint8).reshape(-1))

def test_gpu_encoder_to_gpu_decoder(B):
    """encode on the GPU, decode on the GPU, one frame per decode call (state persists across calls)"""
    for cid in ("c2_gop_range_24sl", "c3_422p10_ctx1", "c4_gbrp14_30sl"):
        case = [c for c in CASES if c[0] == cid][0]
        _, w, h, fmt, opts, kind, n = case
        frames = make_frames(case)
        o = dict(opts); o["g"] = o.pop("gop")
        enc = B.FFV1Encoder(w, h, fmt, max_batch_frames=8, **o)
        pkts = enc.encode_batch(frames)
        dec = B.FFV1Decoder(w, h, enc.extradata, max_batch_frames=8)
        for i, f in enumerate(frames):
            out, key, dmg = dec.decode(pkts[i][0])
            assert key == pkts[i][1] and dmg == 0
            assert np.array_equal(out, np.ascontiguousarray(f).view(np.uint8).reshape(-1)), "%s frame %d" % (cid, i)

def test_damaged_slice_is_flagged_and_concealed(B):
    case = [c for c in CASES if c[0] == "c1_cif_intra"][0]
    cid, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    bad = bytearray(pkts[1])
    pos = len(bad) // 2
    bad[pos] ^= 0x55                                                      # lands in one of the middle slices
    # which slice holds that byte: walk the footers back to front (ffv1dec.c:948-961); trailer = len24 + 0x00 + crc32
    end, omask = len(bad), 0
    for si in range(3, -1, -1):
        size = int.from_bytes(pkts[1][end - 8:end - 5], "big") + 8
        if end - size <= pos < end:
            omask = 1 << si
        end -= size
    assert end == 0 and omask
    d = B.FFV1Decoder(w, h, extradata, max_batch_frames=4)
    got = d.decode_batch([pkts[0], bytes(bad), pkts[2]])
    assert got[0][2] == 0 and got[2][2] == 0
    assert got[1][2] == omask
    # concealment (ffv1dec.c:998-1021): the damaged slice shows the previous frame's pixels, the others decode normally
    f0 = got[0][0][:w * h].reshape(h, w); f1 = got[1][0][:w * h].reshape(h, w)
    good = ref[1][0][:w * h].reshape(h, w)
    for si in range(4):
        ys, xs = slice(h // 2 * (si // 2), h // 2 * (si // 2 + 1)), slice(w // 2 * (si % 2), w // 2 * (si % 2 + 1))
        if omask >> si & 1:
            assert np.array_equal(f1[ys, xs], f0[ys, xs])
        else:
            assert np.array_equal(f1[ys, xs], good[ys, xs])

def test_non_keyframe_first_is_refused(B):
    case = [c for c in CASES if c[0] == "c2_gop_range_24sl"][0]
    cid, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    d = B.FFV1Decoder(w, h, extradata)
    with pytest.raises(B.FFV1Error) as e:
        d.decode(pkts[1])
    assert e.value.code == -1094995529      # AVERROR_INVALIDDATA (ffv1dec.c:930-935)

@pytest.mark.parametrize("env", [{"FFV1B200_DEC_SMEM": "0", "FFV1B200_DEC_PIPE": "0"}, {"FFV1B200_DEC_PIPE": "0", "FFV1B200_DEC_MINB": "12"},
                                 {"FFV1B200_DEC_SMEM": "0", "FFV1B200_DEC_MINB": "12", "FFV1B200_DEC_PIPE": "0"}, {"FFV1B200_DEC_PIPE": "0"},
                                 {"FFV1B200_DEC_PIPE": "1"}],
                         ids=["models_global", "dense_regs", "models_global_dense_regs", "one_warp_per_chain", "two_warps_per_chain"])
def test_large_batch_kernel_variants(B, env, monkeypatch):
    """launch_decode picks the kernel form by the number of chains in the batch: two warps per chain (luma a frame ahead
    of chroma) for small batches, one warp per chain beyond, models in shared or global memory, two register
    allocations; every form is forced here and must decode the same bytes"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for cid in ("c2_gop_range_24sl", "c1_cif_intra", "c3_422p10_ctx1", "fate_ffv1_golomb", "fate_v3_444p16", "yuva420p", "gray16", "v0_range"):
        sel = [c for c in CASES if c[0] == cid]
        if not sel:
            continue
        case = sel[0]
        _, w, h, fmt, opts, kind, n = case
        frames, extradata, pkts, ref = oracle_stream(case)
        got = B.FFV1Decoder(w, h, extradata, max_batch_frames=len(pkts)).decode_batch(pkts)
        for i in range(len(pkts)):
            assert got[i][2] == 0
            assert np.array_equal(got[i][0], ref[i][0]), "%s: frame %d differs (%s)" % (cid, i, env)


def _pinned(nbytes):
    import torch
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()

@pytest.mark.parametrize("cid", ["c2_gop_range_24sl", "fate_ffv1_golomb", "c3_422p10_ctx1", "c4_gbrp14_30sl", "yuva420p", "yuv410p_odd", "v0_range"])
def test_pictures_written_straight_into_pinned_memory(B, cid):
    """an output array in pinned (device-mapped) host memory is written by the kernel itself, no copy behind the batch:
    same pictures as through the staged copy, for every layout family"""
    sel = [c for c in CASES if c[0] == cid]
    if not sel:
        pytest.skip("no such case")
    case = sel[0]
    _, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    fb = len(ref[0][0])
    out = _pinned(len(pkts) * fb)
    out[:] = 0xA5                                                         # nothing may be left of this
    got = B.FFV1Decoder(w, h, extradata, max_batch_frames=len(pkts)).decode_batch(pkts, out=out)
    for i in range(len(pkts)):
        assert got[i][2] == 0
        assert np.array_equal(got[i][0], ref[i][0]), "%s: frame %d differs" % (cid, i)

def test_refused_and_damaged_slices_in_pinned_memory(B, monkeypatch):
    """slices the decoder refuses (header garbage) or conceals leave the same bytes in a pinned output array (cleared by
    the kernel) as in a staged one (cleared by a memset)"""
    case = [c for c in CASES if c[0] == "c2_gop_range_24sl"][0]
    _, w, h, fmt, opts, kind, n = case
    frames, extradata, pkts, ref = oracle_stream(case)
    fb = len(ref[0][0])
    bad0 = bytearray(pkts[0])
    # second slice of the first frame: its first bytes hold the slice header
    sizes, end = [], len(bad0)
    for si in range(23, -1, -1):
        size = int.from_bytes(pkts[0][end - 8:end - 5], "big") + 8
        sizes.insert(0, size); end -= size
    assert end == 0
    start1 = sizes[0]
    for k in range(12):
        bad0[start1 + k] = 0xFF                                           # refused or garbage: flagged either way, nothing to conceal from
    bad2 = bytearray(pkts[2]); bad2[len(bad2) // 3] ^= 0x40               # a CRC error in a later frame: concealed from frame 1
    stream = [bytes(bad0), pkts[1], bytes(bad2), pkts[3]]
    res = {}
    for mode in ("staged", "pinned"):
        out = _pinned(len(stream) * fb) if mode == "pinned" else np.empty(len(stream) * fb, np.uint8)
        out[:] = 0x5A
        got = B.FFV1Decoder(w, h, extradata, max_batch_frames=len(stream)).decode_batch(stream, out=out)
        res[mode] = [(g[0].copy(), g[1], g[2]) for g in got]
    assert res["staged"][0][2] != 0 and res["staged"][2][2] != 0
    for a, b in zip(res["staged"], res["pinned"]):
        assert a[1] == b[1] and a[2] == b[2]
        assert np.array_equal(a[0], b[0])
