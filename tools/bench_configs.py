#!/usr/bin/env python3
"""Encode throughput of the BASELINE.json configurations that are not the headline metric, each next to the reference's
own CPU encoder on the same options and the same clip (one JSON line per configuration):

  c3   configs[2]  1920x1080 yuv422p10le, GOP 16, coder 0 requested (the reference forces the range coder above 8 bits),
                   context=1 (large context model: 7563 contexts), 4 slices               -- clip S3
  c4   configs[3]  3840x2160 gbrp14le (the legal stand-in for RGB48), GOP 16, coder=2, 30 slices   -- clip S4
  gr   configs[1]'s clip with coder=0: Golomb-Rice / run mode, the reference's default for 8-bit content

usage: bench_configs.py [c3|c4|gr|p10|uhd|bgr0 ...] [--frames N] [--steps K]
value = frames resident in HBM (CUDA events); e2e = pinned host frames -> host packets through submit_host/collect_async.
The packets of these configurations are compared with the oracle at full size by tests/test_gpu_encode.py
(test_full_size_configs); here the stream is only decoded back by the CUDA decoder (first GOP, bit-exact round trip)."""
import argparse, ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import numpy as np

def clip_s3(n):
    """S3: S2's formulas at 10 bits, 4:2:2, seed 1235 (planes uint16 little endian)"""
    W, H = 1920, 1080
    rng = np.random.default_rng(1235)
    yy, xx = np.mgrid[0:H, 0:W]
    cx, cy = xx[:, ::2], yy[:, ::2]
    out = []
    for k in range(n):
        Y = np.clip(4 * ((0.1 * xx + 0.07 * yy + 1.5 * k) % 256) + rng.normal(0, 8, (H, W)), 0, 1023)
        U = np.clip(512 + 80 * np.sin((cx + 3 * k) / 97) + rng.normal(0, 6, cx.shape), 0, 1023)
        V = np.clip(512 + 80 * np.cos((cy + 2 * k) / 71) + rng.normal(0, 6, cy.shape), 0, 1023)
        out.append(np.concatenate([p.astype("<u2").ravel() for p in (Y, U, V)]).view(np.uint8))
    return np.stack(out)

def clip_p10(n):
    """S2's formulas at 10 bits, 4:2:0, seed 1237 (planes uint16 little endian): 1080p yuv420p10le, the small context model"""
    W, H = 1920, 1080
    rng = np.random.default_rng(1237)
    yy, xx = np.mgrid[0:H, 0:W]
    cx, cy = xx[::2, ::2], yy[::2, ::2]
    out = []
    for k in range(n):
        Y = np.clip(4 * ((0.1 * xx + 0.07 * yy + 1.5 * k) % 256) + rng.normal(0, 8, (H, W)), 0, 1023)
        U = np.clip(512 + 80 * np.sin((cx + 3 * k) / 97) + rng.normal(0, 6, cx.shape), 0, 1023)
        V = np.clip(512 + 80 * np.cos((cy + 2 * k) / 71) + rng.normal(0, 6, cy.shape), 0, 1023)
        out.append(np.concatenate([p.astype("<u2").ravel() for p in (Y, U, V)]).view(np.uint8))
    return np.stack(out)

def clip_uhd(n):
    """S2's formulas at 3840x2160, seed 1238"""
    W, H = 3840, 2160
    rng = np.random.default_rng(1238)
    yy, xx = np.mgrid[0:H, 0:W]
    out = []
    for k in range(n):
        Y = np.clip(((0.1 * xx + 0.07 * yy + 1.5 * k) % 256) + rng.normal(0, 2, (H, W)), 0, 255)
        U = np.clip(128 + 20 * np.sin((xx[::2, ::2] + 3 * k) / 97) + rng.normal(0, 1.5, (H // 2, W // 2)), 0, 255)
        V = np.clip(128 + 20 * np.cos((yy[::2, ::2] + 2 * k) / 71) + rng.normal(0, 1.5, (H // 2, W // 2)), 0, 255)
        out.append(np.concatenate([p.astype(np.uint8).ravel() for p in (Y, U, V)]))
    return np.stack(out)

def clip_bgr0(n):
    """1920x1080 bgr0: three correlated 8-bit channels (G = S2's luma formula, B and R = G + noise), seed 1239"""
    W, H = 1920, 1080
    rng = np.random.default_rng(1239)
    yy, xx = np.mgrid[0:H, 0:W]
    out = []
    for k in range(n):
        G = np.clip(((0.1 * xx + 0.07 * yy + 1.5 * k) % 256) + rng.normal(0, 2, (H, W)), 0, 255)
        B = np.clip(G + rng.normal(0, 1.5, (H, W)), 0, 255)
        R = np.clip(G + rng.normal(0, 1.5, (H, W)), 0, 255)
        px = np.zeros((H, W, 4), np.uint8)
        px[:, :, 0], px[:, :, 1], px[:, :, 2] = B.astype(np.uint8), G.astype(np.uint8), R.astype(np.uint8)
        out.append(px.ravel())
    return np.stack(out)

def clip_s4(n):
    """S4: three correlated 14-bit planes G, B, R (gbrp14le), seed 1236"""
    W, H = 3840, 2160
    rng = np.random.default_rng(1236)
    yy, xx = np.mgrid[0:H, 0:W]
    out = []
    for k in range(n):
        G = np.clip(64 * ((0.1 * xx + 0.07 * yy + 1.5 * k) % 256) + rng.normal(0, 32, (H, W)), 0, 16383)
        B = np.clip(G + rng.normal(0, 24, (H, W)), 0, 16383)
        R = np.clip(G + rng.normal(0, 24, (H, W)), 0, 16383)
        out.append(np.concatenate([p.astype("<u2").ravel() for p in (G, B, R)]).view(np.uint8))
    return np.stack(out)

def clip_s2(n):
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    return b.s2_clip(n)

CONFIGS = {
    "c3": dict(w=1920, h=1080, fmt="yuv422p10le", opts=dict(level=3, coder=0, context=1), clip=clip_s3, nclip=16, frames=512,
               what="BASELINE configs[2]: 1080p yuv422p10, GOP 16, coder 0 requested (range coder forced), context=1, 4 slices"),
    "c4": dict(w=3840, h=2160, fmt="gbrp14le", opts=dict(level=3, coder=2, context=0, slices=30), clip=clip_s4, nclip=4, frames=64,
               what="BASELINE configs[3]: 2160p gbrp14le (RCT, 15-bit residuals), GOP 16, coder=2, 30 slices"),
    "p10": dict(w=1920, h=1080, fmt="yuv420p10le", opts=dict(level=3, coder=1, context=0, slices=24), clip=clip_p10, nclip=16, frames=512,
                what="1080p yuv420p10le, GOP 16, range coder, context=0 (666 contexts), 24 slices: the 10-bit sibling of configs[1]"),
    "uhd": dict(w=3840, h=2160, fmt="yuv420p", opts=dict(level=3, coder=1, context=0, slices=30), clip=clip_uhd, nclip=8, frames=512,
                what="2160p yuv420p8, GOP 16, range coder, context=0, 30 slices (640-sample slices)"),
    "bgr0": dict(w=1920, h=1080, fmt="bgr0", opts=dict(level=3, coder=1, context=0, slices=24), clip=clip_bgr0, nclip=16, frames=512,
                what="1080p bgr0 (packed 8-bit RGB through the RCT, 9-bit residuals), GOP 16, range coder, 24 slices"),
    "gr": dict(w=1920, h=1080, fmt="yuv420p", opts=dict(level=3, coder=0, context=0, slices=24), clip=clip_s2, nclip=32, frames=1024,
               what="configs[1]'s clip with coder=0: Golomb-Rice / run mode, 24 slices"),
}

def run(name, frames=None, steps=3, ref_frames=16):
    import torch, ffv1_b200
    from ffv1_b200.codec import plane_shapes
    cfg = CONFIGS[name]
    W, H, FMT, opts = cfg["w"], cfg["h"], cfg["fmt"], cfg["opts"]
    GOP = 16
    n = max(GOP, (frames or cfg["frames"]) // GOP * GOP)
    dev = torch.device("cuda", 0)
    clip = cfg["clip"](cfg["nclip"])
    fb = clip.shape[1]
    assert fb == ffv1_b200.frame_bytes(FMT, W, H)
    clip_dev = torch.from_numpy(clip).to(dev)
    frames_dev = clip_dev[torch.arange(n, device=dev) % len(clip)].contiguous()
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, max_batch_frames=n, **opts)
    out_cap = n * (fb + 65536)
    out_dev = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    shapes = plane_shapes(FMT, W, H)
    pl, ls = [], []
    for f in range(n):
        off = frames_dev.data_ptr() + f * fb
        for i in range(4):
            if i < len(shapes):
                pl.append(off); ls.append(shapes[i][1]); off += shapes[i][0] * shapes[i][1]
            else:
                pl.append(0); ls.append(0)
    pl = (ctypes.c_void_p * (4 * n))(*pl); ls = (ctypes.c_int * (4 * n))(*ls)
    stream = torch.cuda.Stream(device=dev)
    sh = ctypes.c_void_p(stream.cuda_stream)
    pk = enc.encode_device(pl, ls, out_dev.data_ptr(), out_cap, n, stream=sh)          # warm-up (scratch areas settle)
    pk = enc.encode_device(pl, ls, out_dev.data_ptr(), out_cap, n, stream=sh)
    s0 = enc.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            pk = enc.encode_device(pl, ls, out_dev.data_ptr(), out_cap, n, stream=sh)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    s1 = enc.stats()
    d = {k: getattr(s1, k) - getattr(s0, k) for k, _ in s1._fields_}
    first_gop = [out_dev[pk[i].offset:pk[i].offset + pk[i].size].cpu().numpy().tobytes() for i in range(GOP)]
    dec = ffv1_b200.FFV1Decoder(W, H, enc.extradata, max_batch_frames=GOP)
    got = dec.decode_batch(first_gop)
    for i in range(GOP):
        assert np.array_equal(got[i][0], clip[i % len(clip)]), "frame %d does not round-trip" % i
    dec.close()
    # host path
    host_in = torch.empty((n, fb), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(frames_dev)
    enc.close()
    del frames_dev, out_dev
    torch.cuda.empty_cache()
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, max_batch_frames=n, first_picture_number=0, **opts)
    host_out = [torch.empty(out_cap, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    table = enc.prepare([host_in[i].numpy() for i in range(n)])
    enc.submit(table); enc.collect(out=host_out[0], copy=False)
    t0 = time.perf_counter()
    enc.submit(table)
    for i in range(steps - 1):
        enc.submit(table)
        enc.collect(out=host_out[i & 1], copy=False, wait_bytes=False)
    pk2 = enc.collect(out=host_out[(steps - 1) & 1], copy=False)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    enc.close()
    line = {"config": name, "workload": cfg["what"], "value": n * steps / (ms * 1e-3), "unit": "frames/s", "frames_per_step": n,
            "steps": steps, "ms_per_step": ms / steps,
            "kernels_ms_per_step": {"pixel": d["ms_pixel_kernel"] / steps, "state_replay": d["ms_model_kernel"] / steps,
                                    "coder": d["ms_coder_kernel"] / steps, "pack_crc": d["ms_pack_kernel"] / steps},
            "packet_bytes_per_frame": sum(p.size for p in pk) / n, "round_trip": "first GOP bit-exact through the CUDA decoder",
            "e2e": {"value": n * steps / dt, "unit": "frames/s", "h2d_bytes_per_step": n * fb, "d2h_bytes_per_step": int(sum(p.size for p in pk2))}}
    if ref_frames:
        from oracle import ffv1_ref
        if ffv1_ref.available():
            threads = min(os.cpu_count() or 1, 24)
            r = ffv1_ref.Encoder(W, H, FMT, gop=GOP, threads=threads, **opts)
            fr = [clip[i % len(clip)] for i in range(ref_frames)]
            r.encode(fr[0])
            t0 = time.perf_counter()
            for f in fr:
                r.encode(f)
            dtr = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ref_frames / dtr, "unit": "frames/s", "cores": threads, "kind": "reference",
                                    "sample": "%d frames, unmodified ffv1enc.c (oracle/_ref), slice threads" % ref_frames}
            line["ratio_e2e"] = line["e2e"]["value"] / line["cpu_baseline"]["value"]
    return line

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c3", "c4", "gr"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    for c in a.configs:
        print(json.dumps(run(c, a.frames or None, a.steps)), flush=True)
