#!/usr/bin/env python3
"""bench.py -- FFV1 encode hot path on B200 (BASELINE.json: "1080p FFV1 P-frame encode fps bit-exact ...; pred-kernel
HBM GB/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]        our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference ...                                     the reference's own CPU encoder (oracle/_ref)
  torchrun --nproc-per-node N ... bench.py --gpus N ...                    one rank per GPU, GOP-aligned frame ranges

Workload (BASELINE.json configs[1]): 1920x1080 yuv420p 8-bit synthetic "camera noise" clip, FFV1 level 3, GOP 16
(state-carry-over non-keyframes = the reference's P-frames), coder=1 (range coder, custom table), context=0,
24 slices, slice CRCs.  A step = one batch of B frames per GPU through the whole encode path (per-pixel pass, state
replay, range coder, packet assembly).  value = frames/s with frames resident in HBM; e2e = the same through the
host-buffer C-ABI call (pinned host frames in, packets out to host memory, copies inside the timed region).
"""
import argparse, json, os, statistics, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))

W, H, FMT = 1920, 1080, "yuv420p"
OPTS = dict(level=3, coder=1, context=0, slices=24)
GOP = 16
FRAME_BYTES = W * H * 3 // 2
SAMPLES = W * H * 3 // 2
ALGO_BYTES_PER_SAMPLE = 5          # 1 B read + 4 B (context,diff) record written
# dram__bytes_read.sum + dram__bytes_write.sum of k_pixel_fast per frame, from the ncu --set full capture summarised in
# profiles/r01_k_pixel_fast.txt (6.015 GB + 12.686 GB for a 1024-frame launch; the reads include the two halo rows per
# 32-row item, the 16-byte alignment blocks either side of a 320-byte row and DRAM's 64-byte access granularity)
TRAFFIC_BYTES_PER_FRAME = (6014702000 + 12686427000) / 1024
METRIC = "1080p yuv420p8 FFV1 level-3 GOP-16 encode throughput (bit-exact)"
WORKLOAD = "1080p yuv420p8 synthetic noise clip, FFV1 level 3, GOP 16 (P-frames), coder=1, context=0, 24 slices, slicecrc"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synth_frames_torch(torch, nframes, device, seed):
    """S2 'noisy1080'-style frames (ramps + gaussian noise) generated on the GPU, tightly packed yuv420p"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((nframes, FRAME_BYTES), dtype=torch.uint8, device=device)
    yy, xx = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32), torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    cx, cy = xx[::2, ::2], yy[::2, ::2]
    for n in range(nframes):
        Y = ((0.1 * xx + 0.07 * yy + 1.5 * n) % 256) + 2.0 * torch.randn((H, W), device=device, generator=g)
        U = 128 + 20 * torch.sin((cx + 3 * n) / 97) + 1.5 * torch.randn(cx.shape, device=device, generator=g)
        V = 128 + 20 * torch.cos((cy + 2 * n) / 71) + 1.5 * torch.randn(cy.shape, device=device, generator=g)
        out[n, :W * H] = Y.clamp(0, 255).to(torch.uint8).reshape(-1)
        out[n, W * H:W * H + W * H // 4] = U.clamp(0, 255).to(torch.uint8).reshape(-1)
        out[n, W * H + W * H // 4:] = V.clamp(0, 255).to(torch.uint8).reshape(-1)
    return out


def synth_frames_numpy(nframes, seed=1234):
    from oracle import synth
    g = synth.Noisy(W, H, FMT, seed)
    return [g.next() for _ in range(nframes)]


def run_reference(args, rank, world):
    """The reference's own CPU encoder (unmodified sources compiled into oracle/_ref), slice-threaded on the host cores."""
    if rank != 0:
        return
    from oracle import ffv1_ref
    if not ffv1_ref.available():
        from oracle import ffv1_oracle
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libffv1ref.so missing (build needs /root/reference)"}))
        return
    threads = min(os.cpu_count() or 1, 24)
    per_step = args.ref_frames
    frames = synth_frames_numpy(per_step)
    enc = ffv1_ref.Encoder(W, H, FMT, gop=GOP, threads=threads, **OPTS)
    nbytes = 0
    for _ in range(args.warmup):
        for f in frames:
            enc.encode(f)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for f in frames:
            nbytes += len(enc.encode(f)[0])
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step, "host_threads": threads},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
                             "sample": "%d frames/step x %d steps, slice threads (pthread_slice.c), unmodified ffv1enc.c" % (per_step, args.steps)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_leg(nframes=48):
    """reference CPU encoder on a bounded sample of the same workload (rank 0, N=1 only)"""
    try:
        from oracle import ffv1_ref
        kind = "reference"
        if ffv1_ref.available():
            threads = min(os.cpu_count() or 1, 24)
            enc = ffv1_ref.Encoder(W, H, FMT, gop=GOP, threads=threads, **OPTS)
        else:
            from oracle import ffv1_oracle
            kind, threads = "port", 1
            nframes = 16
            enc = ffv1_oracle.Encoder(W, H, FMT, gop=GOP, **OPTS)
        frames = synth_frames_numpy(nframes)
        enc.encode(frames[0])      # warm the page cache / thread pool; GOP position restarts below anyway
        t0 = time.perf_counter()
        for f in frames:
            enc.encode(f)
        dt = time.perf_counter() - t0
        return {"value": nframes / dt, "unit": "frames/s", "cores": threads, "kind": kind,
                "sample": "%d frames of the same workload, %s, %d host thread(s)" % (nframes, "oracle/_ref slice-threaded" if kind == "reference" else "oracle C port", threads)}
    except Exception as ex:      # the baseline must never take the GPU number down
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2048, help="frames per step per GPU (multiple of the GOP size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-frames", type=int, default=32, help="--impl reference: frames per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--coder", type=int, default=None, help="variant runs only: 0 = Golomb-Rice, -2 = default state table (the metric is coder=1)")
    ap.add_argument("--context", type=int, default=None, help="variant runs only: 1 = large context model")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the decoder leg (extra key \"decode\", N=1 only)")
    args = ap.parse_args()
    global WORKLOAD
    if args.coder is not None or args.context is not None:
        if args.coder is not None: OPTS["coder"] = args.coder
        if args.context is not None: OPTS["context"] = args.context
        WORKLOAD = WORKLOAD.replace("coder=1, context=0", "VARIANT coder=%d, context=%d" % (OPTS["coder"], OPTS["context"]))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import ffv1_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FFV1 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.pop("NCCL_DEBUG", None)     # WARN / VERSION / INFO all print a banner; stdout is also redirected below
        dist.init_process_group("nccl", device_id=dev)

    B = max(GOP, args.batch // GOP * GOP)
    steps_total = args.warmup + args.steps
    # GOP-aligned frame ranges: rank r codes pictures [r*steps_total*B, ...) -- no data is exchanged between ranks
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=B,
                                first_picture_number=rank * 2 * steps_total * B, **OPTS)
    frames_dev = synth_frames_torch(torch, B, dev, 1234 + rank)
    out_cap = B * (FRAME_BYTES // 2 + 65536)
    out_dev = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    base = frames_dev.data_ptr()
    planes, ls = [], []
    for f in range(B):
        p0 = base + f * FRAME_BYTES
        planes += [p0, p0 + W * H, p0 + W * H + W * H // 4, 0]
        ls += [W, W // 2, W // 2, 0]
    import ctypes
    planes = (ctypes.c_void_p * (4 * B))(*planes)
    ls = (ctypes.c_int * (4 * B))(*ls)
    stream = torch.cuda.Stream(device=dev)
    sh = ctypes_ptr(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ("value")
    for _ in range(args.warmup):
        pk = enc.encode_device(planes, ls, out_dev.data_ptr(), out_cap, B, stream=sh)
    s0 = enc.stats()
    st0 = {k: getattr(s0, k) for k, _ in s0._fields_}
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            pk = enc.encode_device(planes, ls, out_dev.data_ptr(), out_cap, B, stream=sh)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    s1 = enc.stats()
    st1 = {k: getattr(s1, k) for k, _ in s1._fields_}
    d = {k: st1[k] - st0[k] for k in st1}
    pkt_bytes_step = sum(p.size for p in pk)

    # ---------------- end to end through the host-buffer C-ABI calls ("e2e"): pinned host frames in, packets out to
    #                  pinned host memory; submit/collect keeps two batches in flight (copies overlap kernels)
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty((B, FRAME_BYTES), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(frames_dev)
        host_frames = [host_in[i].numpy() for i in range(B)]
        enc.close()
        del enc
        torch.cuda.empty_cache()
        enc2 = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=B,
                                     first_picture_number=(rank * 2 + 1) * steps_total * B, **OPTS)
        host_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True).numpy()
        table = enc2.prepare(host_frames)
        for _ in range(min(args.warmup, 2)):
            enc2.submit(table)
            enc2.collect(out=host_out, copy=False)
        barrier()
        t0 = time.perf_counter()
        enc2.submit(table)
        for _ in range(args.steps - 1):
            enc2.submit(table)
            pk2 = enc2.collect(out=host_out, copy=False)
        pk2 = enc2.collect(out=host_out, copy=False)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": B * FRAME_BYTES,
               "d2h_bytes_per_step": int(sum(p.size for p in pk2)) + 12 * B + 72,
               "note": "ffv1b200_enc_submit_host/_collect: pinned host frames -> packets in pinned host memory, two batches "
                       "in flight, wall clock over all steps incl. every copy, max over ranks"}
        del enc2

    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = measured_peak()
        px_ms = d["ms_pixel_kernel"] / args.steps              # one k_pixel launch per step
        algo = ALGO_BYTES_PER_SAMPLE * SAMPLES * B             # bytes per launch
        achieved = algo / (px_ms * 1e-3) / 1e9 if px_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "input_bytes_per_step_per_gpu": B * FRAME_BYTES,
                       "l2_policy": "input (%.0f MB) and intermediate streams are far larger than the 126 MB L2" % (B * FRAME_BYTES / 1e6),
                       "parallelism": "GOP-aligned frame ranges per GPU, no collective", "packet_bytes_per_frame": pkt_bytes_step / B},
            "roofline": {"kernel": "k_pixel_fast (prediction/context/residual pass)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": TRAFFIC_BYTES_PER_FRAME * B,
                         "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_k_pixel_fast.txt, scaled to this batch",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo, "launch_ms": px_ms},
            "kernels_ms_per_step": {"pixel": px_ms, "state_replay": d["ms_model_kernel"] / args.steps,
                                    "range_coder": d["ms_coder_kernel"] / args.steps, "pack_crc": d["ms_pack_kernel"] / args.steps},
            "coder": {"binary_decisions_per_frame": d["decisions"] / max(1, d["frames"]),
                      "decisions_per_s_state_replay": d["decisions"] / max(1e-9, d["ms_model_kernel"] * 1e-3),
                      "decisions_per_s_range_coder": d["decisions"] / max(1e-9, d["ms_coder_kernel"] * 1e-3)},
            "gpu_launches": d["kernel_launches"], "retries": d["retries"],
            "clocks": clocks, "e2e": e2e,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        if world == 1 and not args.no_decode:
            try:                                     # decoder of the same stream (extra information, not the metric)
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_decode
                del frames_dev, out_dev
                torch.cuda.empty_cache()
                line["decode"] = bench_decode.run(512, 512, 1, 0 if args.no_cpu_baseline else 33)
            except Exception as ex:
                line["decode"] = {"value": None, "note": repr(ex)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ctypes_ptr(v):
    import ctypes
    return ctypes.c_void_p(v)


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: everything libraries write to fd 1 meanwhile (NCCL banners, ...) goes to stderr
    sys.stdout.flush()
    _saved = os.dup(1)
    os.dup2(2, 1)
    _real_print = print
    _lines = []
    def print(*a, **k):                      # noqa: A001 -- the JSON line is held back until fd 1 is restored
        _lines.append(" ".join(str(x) for x in a))
    try:
        main()
    finally:
        sys.stdout.flush()
        os.dup2(_saved, 1)
        os.close(_saved)
        for ln in _lines:
            _real_print(ln, flush=True)
