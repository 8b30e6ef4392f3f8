#!/usr/bin/env python3
"""bench.py -- FFV1 encode hot path on B200 (BASELINE.json: "1080p FFV1 P-frame encode fps bit-exact ...; pred-kernel
HBM GB/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]        our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference ...                                     the reference's own CPU encoder (oracle/_ref)
  torchrun --nproc-per-node N ... bench.py --gpus N ...                    one rank per GPU, GOP-aligned frame ranges
  python bench.py --frames-total 2400 [--gpus N]                           BASELINE configs[4]: a fixed 2400-frame job,
                                                                           GOP-partitioned over the GPUs, encode + decode

Workload (BASELINE.json configs[1]): 1920x1080 yuv420p 8-bit synthetic "camera noise" clip S2 (32 frames,
numpy seed 1234, tiled in time; with GOP 16 every 32-frame period codes to the same packets), FFV1 level 3, GOP 16
(state-carry-over non-keyframes = the reference's P-frames), coder=1 (range coder, custom table), context=0, 24 slices,
slice CRCs.  BOTH arms encode these bytes.  A step = one batch of B frames per GPU through the whole encode path
(per-pixel pass, state replay, range coder, packet assembly).  value = frames/s with frames resident in HBM; e2e = the
same through the host-buffer C-ABI calls (pinned host frames in, packets out to host memory, copies inside the timed
region).  After the timed loops the packets of the LAST batch of each leg are compared (size + MD5) with the reference
build's own packets (tests/golden/ref_packets.json, written by tests/golden/make_golden.py from oracle/_ref): the line
carries "parity_checked".
"""
import argparse, ctypes, hashlib, json, os, statistics, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))

W, H, FMT = 1920, 1080, "yuv420p"
OPTS = dict(level=3, coder=1, context=0, slices=24)
GOP = 16
CLIP = 32                          # frames of the S2 clip (two GOPs); frame i of any stream = clip frame i % 32
FRAME_BYTES = W * H * 3 // 2
SAMPLES = W * H * 3 // 2
ALGO_BYTES_PER_SAMPLE = 5          # 1 B read + 4 B (context,diff) record written
# dram__bytes_read.sum + dram__bytes_write.sum of k_pixel_fast per frame, from the ncu --set full capture summarised in
# profiles/r01_k_pixel_fast.txt (6.015 GB + 12.686 GB for a 1024-frame launch; the reads include the two halo rows per
# 32-row item, the 16-byte alignment blocks either side of a 320-byte row and DRAM's 64-byte access granularity)
TRAFFIC_BYTES_PER_FRAME = (6014702000 + 12686427000) / 1024
METRIC = "1080p yuv420p8 FFV1 level-3 GOP-16 encode throughput (bit-exact)"
WORKLOAD = "1080p yuv420p8 synthetic noise clip, FFV1 level 3, GOP 16 (P-frames), coder=1, context=0, 24 slices, slicecrc"
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_packets.json")


def config_dict():
    """identical for both arms: it names the workload, nothing about how an arm runs it"""
    return {"workload": WORKLOAD,
            "clip": "S2 noisy1080 (SURVEY.md 8(d)): numpy default_rng(1234), %d frames, tiled in time" % CLIP,
            "l2_policy": "a step's input (>= 99.5 MB per 32 frames, 6.4 GB at the default batch) and intermediates are far "
                         "larger than the 126 MB L2: no flush needed",
            "parallelism": "GOP-aligned frame ranges per GPU, no collective"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def s2_clip(nframes=CLIP, seed=1234):
    """S2 'noisy1080' exactly as SURVEY.md 8(d) states it (draw order Y, U, V per frame); [nframes, FRAME_BYTES] uint8"""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    cx, cy = xx[::2, ::2], yy[::2, ::2]
    out = np.empty((nframes, FRAME_BYTES), np.uint8)
    for n in range(nframes):
        Y = np.clip(((0.1 * xx + 0.07 * yy + 1.5 * n) % 256) + rng.normal(0, 2, (H, W)), 0, 255).astype(np.uint8)
        U = np.clip(128 + 20 * np.sin((cx + 3 * n) / 97) + rng.normal(0, 1.5, cx.shape), 0, 255).astype(np.uint8)
        V = np.clip(128 + 20 * np.cos((cy + 2 * n) / 71) + rng.normal(0, 1.5, cy.shape), 0, 255).astype(np.uint8)
        out[n, :W * H] = Y.ravel()
        out[n, W * H:W * H + W * H // 4] = U.ravel()
        out[n, W * H + W * H // 4:] = V.ravel()
    return out


def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)["s2_noisy1080_c2"]


def check_packets(gold, clip_ok, get_packet, indices, first_frame=0):
    """size, MD5 and key flag of the given packets against the reference build's (frame i <-> golden packet i % 32)"""
    if not clip_ok or OPTS != dict(level=3, coder=1, context=0, slices=24):
        return None
    for i in indices:
        data, key = get_packet(i)
        exp = gold["packets"][(first_frame + i) % CLIP]
        got = [len(data), hashlib.md5(data).hexdigest(), int(key)]
        if got != exp:
            raise SystemExit("PARITY FAILURE: packet %d is %r, the reference encoder gives %r" % (i, got, exp))
    return len(indices)


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


# ------------------------------------------------------------------------------------------------ reference arm
def time_reference(nframes_per_step, steps, warmup):
    """the UNMODIFIED reference encoder (oracle/_ref, compiled from /root/reference by oracle/Makefile) on the S2 clip,
    slice-threaded on the host cores (pthread_slice.c); the oracle C port only if that library did not travel"""
    from oracle import ffv1_ref
    clip = s2_clip(min(nframes_per_step, CLIP))
    frames = [clip[i % len(clip)] for i in range(nframes_per_step)]
    if ffv1_ref.available():
        kind, threads = "reference", min(os.cpu_count() or 1, 24)
        enc = ffv1_ref.Encoder(W, H, FMT, gop=GOP, threads=threads, **OPTS)
        what = "unmodified ffv1enc.c (oracle/_ref), %d slice threads" % threads
    else:
        from oracle import ffv1_oracle
        kind, threads = "port", 1
        enc = ffv1_oracle.Encoder(W, H, FMT, gop=GOP, **OPTS)
        what = "oracle C port, 1 thread (oracle/_ref did not travel)"
    gold = golden()
    clip_ok = hashlib.md5(clip.tobytes()).hexdigest() == gold["input_md5"] if len(clip) == CLIP else False
    for _ in range(warmup):
        for f in frames:
            enc.encode(f)
    pk = []
    t0 = time.perf_counter()
    for _ in range(steps):
        pk = [enc.encode(f) for f in frames]
    dt = time.perf_counter() - t0
    # (frames per step is a multiple of the GOP size, so every step starts on a keyframe like the golden stream)
    parity = check_packets(gold, clip_ok and nframes_per_step % GOP == 0, lambda i: pk[i], range(min(len(pk), CLIP)))
    fps = nframes_per_step * steps / dt
    return fps, dt, {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                     "sample": "%d frames/step x %d steps of the same clip, %s" % (nframes_per_step, steps, what),
                     "parity_checked": bool(parity)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step = max(GOP, args.ref_frames // GOP * GOP)
    fps, dt, cb = time_reference(per_step, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(),
            "run": {"frames_per_step": per_step, "host_threads": cb["cores"],
                    "note": "ONE process with up to 24 slice threads whatever --gpus says (slice threading is all the "
                            "reference offers inside one encoder; 24 slices cap it): at N GPUs the ratio compares N GPUs "
                            "with these threads"},
            "cpu_baseline": cb, "parity_checked": cb["parity_checked"],
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2048, help="frames per step per GPU (multiple of 32)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-frames", type=int, default=32, help="--impl reference: frames per step")
    ap.add_argument("--frames-total", type=int, default=0, help="BASELINE configs[4]: a fixed job of this many frames split "
                    "into GOP-aligned ranges over the GPUs (strong scaling), encode and decode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--coder", type=int, default=None, help="variant runs only: 0 = Golomb-Rice, -2 = default state table (the metric is coder=1)")
    ap.add_argument("--context", type=int, default=None, help="variant runs only: 1 = large context model")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the decoder leg (extra key \"decode\", N=1 only)")
    ap.add_argument("--no-avcodec", action="store_true", help="skip the leg through the AVCodec shim (extra key \"e2e_avcodec\")")
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
    # the rank's thread and its pinned staging buffers go to the NUMA node the GPU hangs off (before anything is pinned)
    numa_node = ffv1_b200.bind_thread_to_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gold = golden()
    clip = s2_clip()
    clip_ok = hashlib.md5(clip.tobytes()).hexdigest() == gold["input_md5"]
    clip_dev = torch.from_numpy(clip).to(dev)

    if args.frames_total:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_job
        line = bench_job.run(args, rank, local_rank, world, dev, dist, barrier, max_over_ranks, clip, clip_dev, clip_ok, gold,
                             config_dict(), METRIC, OPTS, numa_node, ClockSampler)
        if line is not None:
            print(json.dumps(line))
        if dist is not None:
            dist.destroy_process_group()
        return

    B = max(CLIP, args.batch // CLIP * CLIP)
    steps_total = args.warmup + args.steps
    # GOP-aligned frame ranges: rank r codes pictures [r*2*steps_total*B, ...) -- no data is exchanged between ranks
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=B,
                                first_picture_number=rank * 2 * steps_total * B, **OPTS)
    frames_dev = clip_dev[torch.arange(B, device=dev) % CLIP].contiguous()          # [B, FRAME_BYTES]
    out_cap = B * (FRAME_BYTES // 2 + 65536)
    out_dev = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    base = frames_dev.data_ptr()
    planes, ls = [], []
    for f in range(B):
        p0 = base + f * FRAME_BYTES
        planes += [p0, p0 + W * H, p0 + W * H + W * H // 4, 0]
        ls += [W, W // 2, W // 2, 0]
    planes = (ctypes.c_void_p * (4 * B))(*planes)
    ls = (ctypes.c_int * (4 * B))(*ls)
    stream = torch.cuda.Stream(device=dev)
    sh = ctypes.c_void_p(stream.cuda_stream)

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
    # parity of the timed path: the first and the last 32 packets of the last timed batch against the reference's own
    tail = list(range(CLIP)) + list(range(B - CLIP, B))
    def dev_packet(i):
        return out_dev[pk[i].offset:pk[i].offset + pk[i].size].cpu().numpy().tobytes(), pk[i].flags & 1
    parity_value = check_packets(gold, clip_ok, dev_packet, tail)

    # ---------------- end to end through the host-buffer C-ABI calls ("e2e"): pinned host frames in, packets out to
    #                  pinned host memory; submit / collect_async keep two batches in flight: the frames of batch k+2 go
    #                  to the device while batch k+1 is coded and the packets of batch k go to the host
    e2e, parity_e2e = None, None
    if not args.no_e2e:
        host_in = torch.empty((B, FRAME_BYTES), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(frames_dev)
        host_frames = [host_in[i].numpy() for i in range(B)]
        enc.close()
        del enc
        torch.cuda.empty_cache()
        enc2 = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=B,
                                     first_picture_number=(rank * 2 + 1) * steps_total * B, **OPTS)
        host_out = [torch.empty(out_cap, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
        table = enc2.prepare(host_frames)
        for i in range(min(args.warmup, 2)):
            enc2.submit(table)
            enc2.collect(out=host_out[i & 1], copy=False)
        barrier()
        se0 = enc2.stats()
        t0 = time.perf_counter()
        enc2.submit(table)
        for i in range(args.steps - 1):
            enc2.submit(table)
            pk2 = enc2.collect(out=host_out[i & 1], copy=False, wait_bytes=False)
        last = host_out[(args.steps - 1) & 1]
        pk2 = enc2.collect(out=last, copy=False)
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        parity_e2e = check_packets(gold, clip_ok, lambda i: (last[pk2[i].offset:pk2[i].offset + pk2[i].size].tobytes(), pk2[i].flags & 1), tail)
        se1 = enc2.stats()
        d2h_step = int(sum(p.size for p in pk2)) + 12 * B + 72
        e2e = {"value": world * B * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": B * FRAME_BYTES,
               "d2h_bytes_per_step": d2h_step, "numa_node": numa_node,
               "kernels_ms_per_step": {k: (getattr(se1, "ms_" + k + "_kernel") - getattr(se0, "ms_" + k + "_kernel")) / args.steps
                                       for k in ("pixel", "model", "coder", "pack")},
               "note": "ffv1b200_enc_submit_host / _collect_async: pinned host frames -> packets in pinned host memory, two "
                       "batches in flight (H2D of batch k+2, kernels of k+1 and D2H of k overlap), wall clock over all steps "
                       "incl. every copy, max over ranks"}
        del enc2
        # ---- the copies alone (same pinned buffers, same sizes, no kernel): the ceiling the link sets for e2e
        cin, cout = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        npk = int(sum(p.size for p in pk2))
        ho = torch.from_numpy(host_out[0])
        def copies(n):
            for _ in range(n):
                with torch.cuda.stream(cin):
                    frames_dev.copy_(host_in, non_blocking=True)
                with torch.cuda.stream(cout):
                    ho[:npk].copy_(out_dev[:npk], non_blocking=True)
            torch.cuda.synchronize(dev)
        copies(1)
        barrier()
        t0 = time.perf_counter()
        copies(max(2, args.steps // 2))
        dtc = max_over_ranks(time.perf_counter() - t0)
        cfps = world * B * max(2, args.steps // 2) / dtc
        e2e["copy_only"] = {"value": cfps, "unit": "frames/s", "frac": e2e["value"] / cfps,
                            "gb_per_s_per_gpu": {"h2d": B * FRAME_BYTES * max(2, args.steps // 2) / dtc / 1e9,
                                                 "d2h": npk * max(2, args.steps // 2) / dtc / 1e9},
                            "note": "the same H2D and D2H copies, concurrently on two streams, no kernels; frac = e2e / this"}
        del host_in, host_frames, table

    ms = max_over_ranks(ms)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = measured_peak()
        px_ms = d["ms_pixel_kernel"] / args.steps              # one k_pixel launch per step
        algo = ALGO_BYTES_PER_SAMPLE * SAMPLES * B             # bytes per launch
        achieved = algo / (px_ms * 1e-3) / 1e9 if px_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(),
            "run": {"frames_per_step_per_gpu": B, "input_bytes_per_step_per_gpu": B * FRAME_BYTES,
                    "packet_bytes_per_frame": pkt_bytes_step / B},
            "parity_checked": bool(parity_value) and (args.no_e2e or bool(parity_e2e)),
            "parity": {"against": "reference build's packets (tests/golden/ref_packets.json: size, MD5, key flag)",
                       "value_leg_packets": parity_value, "e2e_leg_packets": parity_e2e,
                       "which": "first and last 32 packets of the last timed batch of each leg", "input_md5_ok": clip_ok},
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
        del frames_dev, out_dev
        torch.cuda.empty_cache()
        if world == 1 and not args.no_avcodec and not args.no_e2e:
            try:                                     # the same encode through the AVCodec shim and the reference's libavcodec
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_avcodec
                # as many frames as the e2e leg codes (its batches are 1024 frames: twice the steps)
                line["e2e_avcodec"] = bench_avcodec.run(clip, gold if clip_ok else None, batch=B, opts=OPTS, gop=GOP,
                                                        ref_frames=0 if args.no_cpu_baseline else 64,
                                                        rounds=max(10, 2 * args.steps * B // 2048))
            except Exception as ex:
                line["e2e_avcodec"] = {"value": None, "note": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = time_reference(32, 6, 1)[2]
            except Exception as ex:      # the baseline must never take the GPU number down
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        if world == 1 and not args.no_decode:
            try:                                     # decoder of the same stream (extra information, not the metric)
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_decode
                line["decode"] = bench_decode.run(2048, 2048, 1, 0 if args.no_cpu_baseline else 33)     # 2048-frame batch, like the encoder's step
                small = bench_decode.run(512, 512, 1, 0)                                                # two warps per chain (luma ahead of chroma)
                line["decode"]["batch_512"] = {k: small[k] for k in ("value", "unit", "frames", "batch", "round_trip", "kernel_fps")}
            except Exception as ex:
                line["decode"] = {"value": None, "note": repr(ex)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
