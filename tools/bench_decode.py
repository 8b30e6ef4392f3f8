#!/usr/bin/env python3
"""Decode throughput of the CUDA decoder on the bench workload (1080p yuv420p8, GOP 16, 24 slices, range coder):
encodes N synthetic frames with the CUDA encoder, then times ffv1b200_dec_decode_host over the packets (host packets in,
host frames out, copies included) and checks the round trip.  usage: bench_decode.py [nframes] [batch|0] [coder] [cpu_frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import numpy as np
import ffv1_b200
from oracle import synth

def run(n=256, batch=None, coder=1, cpu_frames=0):
    batch = batch or n
    W, H, FMT = 1920, 1080, "yuv420p"
    g = synth.Noisy(W, H, FMT, 1234)
    base = [g.next() for _ in range(16)]
    frames = [base[i % 16] for i in range(n)]
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=16, level=3, coder=coder, context=0, slices=24, max_batch_frames=min(n, 256))
    pkts = [bytes(p) for p, _ in enc.encode_batch(frames)]
    extradata = enc.extradata
    enc.close()
    dec = ffv1_b200.FFV1Decoder(W, H, extradata, max_batch_frames=batch)
    out = dec.decode_batch(pkts)            # warm-up (+ correctness)
    for i in (0, 1, n // 2, n - 1):
        assert np.array_equal(out[i][0], frames[i]), "frame %d does not round-trip" % i
    host_out, pinned = None, False
    try:                                    # frames land in pinned host memory, like the frames of the encoder's e2e leg
        import torch
        host_out = torch.empty(n * ffv1_b200.frame_bytes(FMT, W, H), dtype=torch.uint8, pin_memory=True).numpy()
        pinned = True
    except Exception:
        pass
    t0 = time.perf_counter()
    out = dec.decode_batch(pkts, out=host_out)
    dt = time.perf_counter() - t0
    for i in (0, n - 1):
        assert np.array_equal(out[i][0], frames[i]), "frame %d does not round-trip" % i
    s = dec.stats()
    st = {k: getattr(s, k) for k, _ in s._fields_}
    dec.close()
    cpu = None
    if cpu_frames:
        # the unmodified reference decoder (oracle/_ref, slice threads) on the first frames of the same stream
        try:
            from oracle import ffv1_ref
            if ffv1_ref.available():
                threads = min(os.cpu_count() or 1, 24)
                rd = ffv1_ref.Decoder(W, H, extradata, threads=threads)
                rd.decode(pkts[0])
                t1 = time.perf_counter()
                for i in range(1, cpu_frames):
                    o, _, _ = rd.decode(pkts[i])
                dtc = time.perf_counter() - t1
                assert np.array_equal(o, frames[cpu_frames - 1]), "reference decoder output differs"
                rd.close()
                cpu = {"value": (cpu_frames - 1) / dtc, "unit": "frames/s", "cores": threads, "kind": "reference",
                       "sample": "%d frames of the same stream, oracle/_ref decoder, slice threads" % (cpu_frames - 1)}
        except Exception as ex:
            cpu = {"value": None, "note": repr(ex)}
    # two decode_batch calls were made (warm-up + timed): the kernel time in the statistics covers both
    return {"value": n / dt, "unit": "frames/s", "frames": n, "batch": batch, "round_trip": "bit-exact",
            "coder": coder, "cpu_baseline": cpu,
            "kernel_fps": 2 * n / (st["ms_decode_kernel"] * 1e-3) if st.get("ms_decode_kernel") else None,
            "note": "ffv1b200_dec_decode_host: host packets in, host frames out (%s: the kernel writes them there itself), packet copy included; "
                    "k_decode = one (GOP, slice) chain per warp, or per warp pair with luma a frame ahead of chroma when the batch has few chains; serial inside a plane like decode_line"
                    % ("pinned host memory" if pinned else "pageable numpy buffers")}

if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) else None,
                         int(sys.argv[3]) if len(sys.argv) > 3 else 1, int(sys.argv[4]) if len(sys.argv) > 4 else 0)))
