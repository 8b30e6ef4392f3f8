"""On-GPU input preparation (-m gpu): ffv1b200_upload_frames (the hwupload_cuda + format-conversion step in front of the
encoder) followed by ffv1b200_enc_encode_cuda must give the packets the oracle produces for the same pixels converted on
the host with numpy, for every source layout, odd sizes included."""
import numpy as np, pytest
from oracle import ffv1_oracle as O

pytestmark = pytest.mark.gpu

def planes_of(fmt, w, h, rng):
    """random source frame: (bytes, [(rows, row bytes)], the frame in the encoder's layout as bytes, encoder pix_fmt)"""
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if fmt == "nv12":
        y = rng.integers(0, 256, (h, w), dtype=np.uint8); uv = rng.integers(0, 256, (ch, cw, 2), dtype=np.uint8)
        src = np.concatenate([y.ravel(), uv.ravel()])
        dst = np.concatenate([y.ravel(), uv[:, :, 0].ravel(), uv[:, :, 1].ravel()])
        return src, [(h, w), (ch, 2 * cw)], dst, "yuv420p"
    if fmt == "p010le":
        y = rng.integers(0, 1024, (h, w)).astype("<u2"); uv = rng.integers(0, 1024, (ch, cw, 2)).astype("<u2")
        src = np.concatenate([(y << 6).ravel(), (uv << 6).ravel()]).astype("<u2").view(np.uint8)
        dst = np.concatenate([y.ravel(), uv[:, :, 0].ravel(), uv[:, :, 1].ravel()]).astype("<u2").view(np.uint8)
        return src, [(h, 2 * w), (ch, 4 * cw)], dst, "yuv420p10le"
    if fmt in ("yuyv422", "uyvy422"):
        y = rng.integers(0, 256, (h, 2 * cw), dtype=np.uint8); u = rng.integers(0, 256, (h, cw), dtype=np.uint8)
        v = rng.integers(0, 256, (h, cw), dtype=np.uint8)
        q = np.empty((h, cw, 4), np.uint8)
        if fmt == "yuyv422":
            q[:, :, 0], q[:, :, 1], q[:, :, 2], q[:, :, 3] = y[:, 0::2], u, y[:, 1::2], v
        else:
            q[:, :, 0], q[:, :, 1], q[:, :, 2], q[:, :, 3] = u, y[:, 0::2], v, y[:, 1::2]
        dst = np.concatenate([y[:, :w].ravel(), u.ravel(), v.ravel()])
        return q.ravel(), [(h, 4 * cw)], dst, "yuv422p"
    if fmt in ("rgb24", "bgr24"):
        p = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b, g, r = (p[:, :, 2], p[:, :, 1], p[:, :, 0]) if fmt == "rgb24" else (p[:, :, 0], p[:, :, 1], p[:, :, 2])
        d = np.zeros((h, w, 4), np.uint8); d[:, :, 0], d[:, :, 1], d[:, :, 2] = b, g, r
        return p.ravel(), [(h, 3 * w)], d.ravel(), "bgr0"
    if fmt == "rgba":
        p = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        d = p[:, :, [2, 1, 0, 3]].copy()
        return p.ravel(), [(h, 4 * w)], d.ravel(), "bgra"
    raise ValueError(fmt)

@pytest.mark.parametrize("size", [(96, 80), (101, 67)], ids=["96x80", "101x67_odd"])
@pytest.mark.parametrize("fmt", ["nv12", "p010le", "yuyv422", "uyvy422", "rgb24", "bgr24", "rgba"])
def test_converted_frames_encode_like_host_converted_ones(fmt, size):
    import ffv1_b200
    w, h = size
    rng = np.random.default_rng(42)
    n = 4
    srcs, dsts = [], []
    for _ in range(n):
        s, shapes, d, efmt = planes_of(fmt, w, h, rng)
        srcs.append(s); dsts.append(d)
    up = ffv1_b200.FFV1Uploader(w, h, fmt, pool_frames=n)
    assert up.pix_fmt == efmt
    opts = dict(level=3, coder=1, slices=4)
    enc = ffv1_b200.FFV1Encoder(w, h, efmt, g=3, max_batch_frames=n, **opts)
    dpl, dls = up.upload(srcs, shapes)
    got = ffv1_b200.encode_cuda(enc, dpl, dls, n)
    o = O.Encoder(w, h, efmt, gop=3, **opts)
    for i in range(n):
        exp, key = o.encode(dsts[i])
        assert got[i][1] == key and got[i][0] == exp, "packet %d differs" % i

def test_upload_only_and_refusals():
    import ffv1_b200
    from oracle import synth
    w, h, fmt = 96, 80, "yuv444p"
    gen = synth.Noisy(w, h, fmt, 3)
    frames = [gen.next() for _ in range(3)]
    up = ffv1_b200.FFV1Uploader(w, h, fmt, pool_frames=3)            # a layout the encoder takes: hwupload alone
    assert up.pix_fmt == fmt
    enc = ffv1_b200.FFV1Encoder(w, h, fmt, g=2, level=3, coder=0, slices=4, max_batch_frames=3)
    dpl, dls = up.upload(frames, ffv1_b200.plane_shapes(fmt, w, h))
    got = ffv1_b200.encode_cuda(enc, dpl, dls, 3)
    o = O.Encoder(w, h, fmt, gop=2, level=3, coder=0, slices=4)
    assert [g for g in got] == [o.encode(f) for f in frames]
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.FFV1Uploader(w, h, "nv12", dst_pix_fmt="yuv444p")
    assert e.value.code == -38
    with pytest.raises(ffv1_b200.FFV1Error):
        ffv1_b200.FFV1Uploader(w, h, "rgb48le")
