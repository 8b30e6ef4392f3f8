#!/usr/bin/env python3
"""The drop-in itself, timed: the bench workload through avcodec_encode_video2 / avcodec_decode_video2 of the REFERENCE's
libavcodec (oracle/_ref/libffv1ref.so, unmodified sources) with the codec "ffv1_b200" (integration/ffv1_b200_avcodec.c ->
libffv1_b200.so) and, through the very same harness loop (oracle/ref_harness.c: ffv1ref_bench_encode / _decode), with the
reference's own "ffv1" on the host cores.  Frames are pageable AVFrames that wrap the clip without a copy (the way ffmpeg.c
hands rawvideo frames on); packets come back as AVPackets.  bench.py adds the result as "e2e_avcodec"."""
import ctypes, hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import numpy as np

SHIM = os.path.join(ROOT, "integration", "_build", "libffv1_b200_avcodec.so")
REF = os.path.join(ROOT, "oracle", "_ref", "libffv1ref.so")
_L = None

def lavc():
    """the reference's libavcodec with the ffv1_b200 codec pair registered (avcodec_register, utils.c:178)"""
    global _L
    if _L is None:
        if not (os.path.exists(SHIM) and os.path.exists(REF)):
            raise RuntimeError("reference build / shim not present (built only where /root/reference exists)")
        import ffv1_b200
        ffv1_b200.lib()
        L = ctypes.CDLL(REF, mode=ctypes.RTLD_GLOBAL)
        shim = ctypes.CDLL(SHIM, mode=ctypes.RTLD_GLOBAL)
        L.ffv1ref_register_codec.argtypes = [ctypes.c_void_p]
        L.ffv1ref_register_codec(ctypes.addressof(ctypes.c_char.in_dll(shim, "ff_ffv1_b200_encoder")))
        L.ffv1ref_register_codec(ctypes.addressof(ctypes.c_char.in_dll(shim, "ff_ffv1_b200_decoder")))
        L.ffv1ref_bench_encode.restype = ctypes.c_double
        L.ffv1ref_bench_encode.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 7 + \
            [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
             ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)]
        L.ffv1ref_bench_decode.restype = ctypes.c_double
        L.ffv1ref_bench_decode.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _L = L
    return _L

def encode(name, clip, w, h, fmt, nframes, gop, opts, threads=1, extra="", nkeep=0):
    """returns (seconds, total packet bytes, [(bytes, key)] of the last nkeep packets)"""
    L = lavc()
    o = dict(level=-1, coder=0, context=0, slices=0, slicecrc=-1); o.update(opts)
    clip = np.ascontiguousarray(clip)
    nclip = clip.shape[0]
    keep = np.empty(max(1, nkeep) * (clip.shape[1] + 65536), np.uint8)
    ksize, kkey, total = (ctypes.c_int * max(1, nkeep))(), (ctypes.c_int * max(1, nkeep))(), ctypes.c_int64()
    dt = L.ffv1ref_bench_encode(name.encode(), w, h, fmt.encode(), gop, o["level"], o["coder"], o["context"], o["slices"], o["slicecrc"],
                                threads, extra.encode(), clip.ctypes.data, nclip, nframes, keep.ctypes.data, keep.nbytes, nkeep,
                                ksize, kkey, ctypes.byref(total))
    if dt < 0:
        raise RuntimeError("ffv1ref_bench_encode(%s) failed: %r" % (name, dt))
    out, off = [], 0
    for k in range(nkeep):
        out.append((keep[off:off + ksize[k]].tobytes(), bool(kkey[k]))); off += ksize[k]
    return dt, total.value, out

def decode(name, w, h, extradata, packets, frame_bytes, threads=1, frame_threads=0, extra=""):
    """returns (seconds, pictures, last picture as uint8 array)"""
    L = lavc()
    n = len(packets)
    bufs = [np.frombuffer(p, np.uint8) for p in packets]
    ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
    sizes = (ctypes.c_int * n)(*[len(p) for p in packets])
    last = np.zeros(frame_bytes, np.uint8)
    nout = ctypes.c_int()
    dt = L.ffv1ref_bench_decode(name.encode(), w, h, extradata, len(extradata), threads, frame_threads, extra.encode(), ptrs, sizes, n,
                                last.ctypes.data, frame_bytes, ctypes.byref(nout))
    if dt < 0:
        raise RuntimeError("ffv1ref_bench_decode(%s) failed: %r" % (name, dt))
    return dt, nout.value, last

def run(clip, gold, batch=1024, opts=None, gop=16, ref_frames=64, copy_threads=8, rounds=10, decode_frames=512):
    W, H, FMT = 1920, 1080, "yuv420p"
    opts = opts or dict(level=3, coder=1, context=0, slices=24)
    batch = max(32, min(batch, 1024) // 32 * 32)
    nframes = batch * rounds
    dt, nbytes, last = encode("ffv1_b200", clip, W, H, FMT, nframes, gop, opts, extra="batch=%d:copy_threads=%d" % (batch, copy_threads), nkeep=32)
    parity = None
    if gold is not None and opts == dict(level=3, coder=1, context=0, slices=24):
        for k, (data, key) in enumerate(last):
            exp = gold["packets"][(nframes - 32 + k) % 32]
            got = [len(data), hashlib.md5(data).hexdigest(), int(key)]
            if got != exp:
                raise SystemExit("PARITY FAILURE through the AVCodec shim: packet %d is %r, the reference gives %r" % (nframes - 32 + k, got, exp))
        parity = len(last)
    res = {"value": nframes / dt, "unit": "frames/s", "frames": nframes, "batch": batch, "copy_threads": copy_threads,
           "packet_bytes": nbytes, "parity_checked": bool(parity),
           "rounds": rounds,
           "note": "avcodec_encode_video2(-c:v ffv1_b200) of the reference's libavcodec: pageable AVFrames in, AVPackets out, "
                   "first frame to last drained packet, i.e. with the pipeline's fill and drain (oracle/ref_harness.c:ffv1ref_bench_encode; "
                   "5 batches: 5.9 k frames/s, 10 batches: 6.5-8.0 k depending on the box)"}
    if ref_frames:
        threads = min(os.cpu_count() or 1, 24)
        dtr, _, _ = encode("ffv1", clip, W, H, FMT, ref_frames, gop, opts, threads=threads)
        res["reference"] = {"value": ref_frames / dtr, "unit": "frames/s", "frames": ref_frames, "cores": threads,
                            "note": "-c:v ffv1 through the same harness loop, slice threads"}
        res["ratio"] = res["value"] / res["reference"]["value"]
    if decode_frames:
        import ffv1_b200
        enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=gop, max_batch_frames=min(decode_frames, 256), **opts)
        pk = [bytes(p) for p, _ in enc.encode_batch([clip[i % len(clip)] for i in range(decode_frames)])]
        ed = enc.extradata
        enc.close()
        fb = clip.shape[1]
        ddt, nout, lastf = decode("ffv1_b200", W, H, ed, pk, fb, extra="batch=%d" % decode_frames)
        assert nout == decode_frames and np.array_equal(lastf, clip[(decode_frames - 1) % len(clip)]), "decode through the shim differs"
        dec = {"value": decode_frames / ddt, "unit": "frames/s", "frames": decode_frames, "batch": decode_frames, "round_trip": "last picture bit-exact",
               "note": "avcodec_decode_video2(-c:v ffv1_b200, batch option, AV_CODEC_CAP_DELAY drain)"}
        ddt1, nout1, _ = decode("ffv1_b200", W, H, ed, pk[:48], fb, extra="batch=1")
        dec["batch_1"] = {"value": nout1 / ddt1, "unit": "frames/s", "frames": nout1, "note": "no delay: one picture per packet"}
        if ref_frames:
            threads = min(os.cpu_count() or 1, 24)
            n = min(len(pk), max(ref_frames, 48))
            rdt, rn, rl = decode("ffv1", W, H, ed, pk[:n], fb, threads=threads)
            assert np.array_equal(rl, clip[(n - 1) % len(clip)])
            dec["reference"] = {"value": rn / rdt, "unit": "frames/s", "frames": rn, "cores": threads, "note": "-c:v ffv1, slice threads, same harness loop"}
            rdt, rn, rl = decode("ffv1", W, H, ed, pk[:n], fb, threads=threads, frame_threads=1)
            dec["reference_frame_threads"] = {"value": rn / rdt, "unit": "frames/s", "frames": rn, "cores": threads}
        res["decode"] = dec
    return res

if __name__ == "__main__":
    import json
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    clip = b.s2_clip()
    print(json.dumps(run(clip, b.golden(), batch=int(sys.argv[1]) if len(sys.argv) > 1 else 1024,
                         copy_threads=int(sys.argv[2]) if len(sys.argv) > 2 else 8)))
