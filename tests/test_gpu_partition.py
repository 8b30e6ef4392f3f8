"""The GOP-aligned partition with the GPU codec itself (-m gpu; the CPU variant in tests/test_partition.py lets the oracle
stand in): every "rank" is an encoder instance opened with first_picture_number = the global index of its range's first
frame and sees only its own frames; the re-interleaved packets must equal the single-stream encode (the oracle's, pinned to
the reference), and every rank's range decodes on its own, starting at its keyframe (ffv1dec.c:930-935)."""
import numpy as np, pytest
from oracle import ffv1_oracle as O, synth

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("fmt,opts", [("yuv420p", dict(level=3, coder=1, context=0, slices=4)),
                                      ("yuv422p10le", dict(level=3, coder=1, context=1, slices=4)),
                                      ("yuv420p", dict(level=3, coder=0, slices=4))], ids=["range", "large_model_10bit", "golomb"])
def test_ranges_coded_independently_equal_the_single_stream(world, fmt, opts):
    import ffv1_b200
    from ffv1_b200 import gop_aligned_ranges, reinterleave
    w, h, gop, nframes = 96, 64, 4, 22
    gen = synth.Noisy(w, h, fmt, 99)
    frames = [gen.next() for _ in range(nframes)]
    single = O.Encoder(w, h, fmt, gop=gop, **opts)
    expect = [single.encode(f) for f in frames]
    per_rank, extradata = [], None
    for start, count in gop_aligned_ranges(nframes, gop, world):
        if not count:
            per_rank.append([])
            continue
        enc = ffv1_b200.FFV1Encoder(w, h, fmt, g=gop, max_batch_frames=5, first_picture_number=start, **opts)
        extradata = enc.extradata
        per_rank.append(enc.encode_batch(frames[start:start + count]))           # batches of 5 against GOPs of 4
        enc.close()
    assert reinterleave(per_rank) == expect
    for (start, count), pk in zip(gop_aligned_ranges(nframes, gop, world), per_rank):
        if not count:
            continue
        dec = ffv1_b200.FFV1Decoder(w, h, extradata, max_batch_frames=8)
        outs = dec.decode_batch([p for p, _ in pk])
        for i in range(count):
            assert np.array_equal(outs[i][0], frames[start + i].view(np.uint8).reshape(-1)), "rank range %d, frame %d" % (start, i)
        dec.close()
