#!/usr/bin/env python3
"""Decode throughput of the CUDA decoder on a BASELINE configuration other than the headline one (see bench_configs.py):
usage: bench_decode_cfg.py c3|c4|gr [frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, ffv1_b200
import bench_configs as C

def run(name, n):
    cfg = C.CONFIGS[name]
    W, H, FMT, opts = cfg["w"], cfg["h"], cfg["fmt"], cfg["opts"]
    clip = cfg["clip"](cfg["nclip"])
    frames = [clip[i % len(clip)] for i in range(n)]
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=16, max_batch_frames=min(n, 64), **opts)
    pkts = [bytes(p) for p, _ in enc.encode_batch(frames)]
    ed = enc.extradata
    enc.close()
    dec = ffv1_b200.FFV1Decoder(W, H, ed, max_batch_frames=n)
    t0 = time.perf_counter()
    out = dec.decode_batch(pkts)
    dt = time.perf_counter() - t0
    for i in (0, n // 2, n - 1):
        assert np.array_equal(out[i][0], frames[i]), "frame %d does not round-trip" % i
    s = dec.stats()
    return {"config": name, "frames": n, "value": n / dt, "unit": "frames/s", "kernel_ms": s.ms_decode_kernel, "round_trip": "bit-exact"}

if __name__ == "__main__":
    print(json.dumps(run(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 64)))
