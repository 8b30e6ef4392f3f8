"""BASELINE.json configs[4]: a FIXED job of T frames (2400 = 150 GOPs of 16) of the bench workload, split into GOP-aligned
frame ranges over the GPUs (ffv1_b200.partition.gop_aligned_ranges: 19,19,19,19,19,19,18,18 GOPs at N=8), each rank
encoding and then decoding only its own range -- strong scaling, no collective on the data path.  Called by
bench.py --frames-total T; returns the JSON line (rank 0) or None.

  encode  value : frames resident in HBM, whole encode path per rank, CUDA events, max over ranks
  encode  e2e   : pinned host frames -> host packets through ffv1b200_enc_submit_host / _collect_async
  decode        : host packets -> pinned host frames through ffv1b200_dec_decode_host (ffv1dec.c:895-1035; a shard starts on
                  a keyframe, ffv1dec.c:930-935)
Checks: every rank's packets of its first and last GOP against the reference build's MD5s, the re-interleaved stream's
packet count, and the decoded frames against the source (all frames of the range)."""
import ctypes, hashlib, os, sys, time

def run(args, rank, local_rank, world, dev, dist, barrier, max_over_ranks, clip, clip_dev, clip_ok, gold, config, metric,
        opts, numa_node, ClockSampler):
    import numpy as np, torch
    import ffv1_b200
    from ffv1_b200.partition import gop_aligned_ranges, reinterleave
    W, H, FMT, GOP, CLIP = 1920, 1080, "yuv420p", 16, len(clip)
    FB = W * H * 3 // 2
    T = args.frames_total
    ranges = gop_aligned_ranges(T, GOP, world)
    start, count = ranges[rank]
    aligned = T % GOP == 0
    # batches of equal size (whole GOPs): a 2400-frame range goes as 1200 + 1200, not 2048 + 352 -- the stages that are bound
    # by the length of a chain (k_rangecode, the replay, the decoder) take as long for 352 frames as for 1200
    nb = max(1, -(-count // max(GOP, args.batch // GOP * GOP)))
    Bmax = max(GOP, -(-(-(-count // nb)) // GOP) * GOP) if count >= GOP else max(1, count)
    steps, warmup = args.steps, args.warmup
    sampler = ClockSampler(local_rank) if rank == 0 else None

    idx = (torch.arange(count, device=dev) + start) % CLIP
    frames_dev = clip_dev[idx].contiguous() if count else torch.empty((0, FB), dtype=torch.uint8, device=dev)
    host_in = torch.empty((max(count, 1), FB), dtype=torch.uint8, pin_memory=True)
    if count:
        host_in[:count].copy_(frames_dev)
    out_cap = max(count, 1) * (FB // 2 + 65536)
    chunks = [(c0, min(Bmax, count - c0)) for c0 in range(0, count, Bmax)]

    def golden_ok(get_packet, n):
        if not (clip_ok and aligned and opts == dict(level=3, coder=1, context=0, slices=24)):
            return None
        which = list(range(min(GOP, n))) + list(range(max(0, n - GOP), n))
        for i in which:
            data, key = get_packet(i)
            exp = gold["packets"][(start + i) % CLIP]
            got = [len(data), hashlib.md5(data).hexdigest(), int(key)]
            if got != exp:
                raise SystemExit("PARITY FAILURE (rank %d): packet %d of the range is %r, the reference encoder gives %r" % (rank, i, got, exp))
        return len(which)

    # ------------------------------------------------------------ encode, frames resident in HBM
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=Bmax, first_picture_number=start, **opts)
    out_dev = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sh = ctypes.c_void_p(stream.cuda_stream)
    tables = []
    for c0, n in chunks:
        base = frames_dev.data_ptr() + c0 * FB
        pl, ls = [], []
        for f in range(n):
            p0 = base + f * FB
            pl += [p0, p0 + W * H, p0 + W * H + W * H // 4, 0]
            ls += [W, W // 2, W // 2, 0]
        tables.append(((ctypes.c_void_p * (4 * n))(*pl), (ctypes.c_int * (4 * n))(*ls), n))

    def device_pass():
        pks, off = [], 0
        for pl, ls, n in tables:
            pk = enc.encode_device(pl, ls, out_dev.data_ptr() + off, out_cap - off, n, stream=sh)
            pks.append((off, pk, n))
            off += sum(p.size for p in pk)
        return pks
    for _ in range(warmup):
        device_pass()
    s0 = enc.stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            pks = device_pass()
        e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) if count else max_over_ranks(0.0)
    s1 = enc.stats()
    flat = [(off + p.offset, p.size, p.flags & 1) for off, pk, n in pks for p in pk]
    par_dev = golden_ok(lambda i: (out_dev[flat[i][0]:flat[i][0] + flat[i][1]].cpu().numpy().tobytes(), flat[i][2]), len(flat))
    enc.close()
    del enc, out_dev
    torch.cuda.empty_cache()

    # ------------------------------------------------------------ encode, host frames -> host packets
    enc = ffv1_b200.FFV1Encoder(W, H, FMT, g=GOP, device=local_rank, max_batch_frames=Bmax,
                                first_picture_number=start + T, **opts)            # (T is GOP-aligned: same GOP phase)
    extradata = enc.extradata
    host_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True).numpy()
    htables = [enc.prepare([host_in[c0 + i].numpy() for i in range(n)]) for c0, n in chunks]

    def host_pass():
        """the whole range through the two-slot pipeline; returns [(offset in host_out, size, key)]"""
        res, off, pending = [], 0, []
        for t in htables:
            enc.submit(t)
            pending.append(t)
            if len(pending) == 2:
                pk = enc.collect(out=host_out[off:], copy=False, wait_bytes=False)
                res += [(off + p.offset, p.size, p.flags & 1) for p in pk]
                off += sum(p.size for p in pk)
                pending.pop(0)
        while pending:
            pk = enc.collect(out=host_out[off:], copy=False, wait_bytes=False)
            res += [(off + p.offset, p.size, p.flags & 1) for p in pk]
            off += sum(p.size for p in pk)
            pending.pop(0)
        enc.sync_output()
        return res
    for _ in range(min(warmup, 2)):
        host_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        flat = host_pass() if count else []
    torch.cuda.synchronize(dev)
    dt_enc = max_over_ranks(time.perf_counter() - t0)
    par_host = golden_ok(lambda i: (host_out[flat[i][0]:flat[i][0] + flat[i][1]].tobytes(), flat[i][2]), len(flat))
    pkt_bytes = sum(s for _, s, _ in flat)
    packets = [host_out[o:o + s].tobytes() for o, s, _ in flat]
    enc.close()
    del enc, frames_dev
    torch.cuda.empty_cache()

    # ------------------------------------------------------------ decode of the rank's own range (starts on a keyframe)
    dec_line = None
    if not args.no_decode:
        dbatch = max(1, min(count, Bmax))
        dec = ffv1_b200.FFV1Decoder(W, H, extradata, device=local_rank, max_batch_frames=dbatch)
        frames_out = torch.empty((max(count, 1), FB), dtype=torch.uint8, pin_memory=True)
        fo = frames_out.numpy().reshape(-1)
        if count:
            dec.decode_batch(packets, out=fo)                                   # warm-up
        barrier()
        t0 = time.perf_counter()
        dsteps = max(1, steps // 2)
        for _ in range(dsteps):
            if count:
                res = dec.decode_batch(packets, out=fo)
        torch.cuda.synchronize(dev)
        dt_dec = max_over_ranks(time.perf_counter() - t0)
        if count:
            assert all(not r[2] for r in res), "a slice failed its CRC"
            assert torch.equal(frames_out[:count], host_in[:count]), "rank %d: decoded frames differ from the source" % rank
        ds = dec.stats()
        dec.close()
        dec_line = {"value": T * dsteps / dt_dec, "unit": "frames/s", "steps": dsteps, "ms_per_step": 1000 * dt_dec / dsteps,
                    "round_trip": "bit-exact (every frame of every range compared with the source)",
                    "batch": dbatch, "kernel_ms_rank0": ds.ms_decode_kernel / (dsteps + 1),
                    "note": "ffv1b200_dec_decode_host per rank on its own GOP-aligned range: host packets in, pinned host frames "
                            "out, copies included, wall clock, max over ranks"}

    # ------------------------------------------------------------ control plane: counts and digests to rank 0
    mine = {"rank": rank, "start": start, "count": count, "packets": len(packets), "packet_bytes": pkt_bytes,
            "stream_md5": hashlib.md5(b"".join(packets)).hexdigest(), "golden_dev": par_dev, "golden_host": par_host}
    gathered = [mine]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return None
    order = reinterleave([[g["rank"]] * g["packets"] for g in gathered])
    assert len(order) == T and order == sorted(order), "re-interleaved stream is not in pts order"
    d = {k: getattr(s1, k) - getattr(s0, k) for k, _ in s1._fields_}
    parity = all(g["golden_dev"] and g["golden_host"] for g in gathered if g["count"])
    return {
        "metric": metric.replace("encode throughput", "encode throughput, fixed %d-frame job (configs[4])" % T),
        "value": T * steps / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": dict(config, job="%d frames = %d GOPs, GOP-aligned ranges %s" % (T, (T + GOP - 1) // GOP, [n for _, n in ranges])),
        "run": {"frames_per_gpu": [n for _, n in ranges], "batch_per_submit": Bmax, "packet_bytes_total": sum(g["packet_bytes"] for g in gathered)},
        "parity_checked": bool(parity),
        "parity": {"against": "reference build's packets (tests/golden/ref_packets.json)", "per_rank": [
            {"rank": g["rank"], "first_frame": g["start"], "frames": g["count"], "golden_packets_value_leg": g["golden_dev"],
             "golden_packets_e2e_leg": g["golden_host"]} for g in gathered],
            "which": "first and last GOP of every rank's range, both legs", "reinterleaved_packets": len(order)},
        "kernels_ms_per_step_rank0": {"pixel": d["ms_pixel_kernel"] / steps, "state_replay": d["ms_model_kernel"] / steps,
                                      "range_coder": d["ms_coder_kernel"] / steps, "pack_crc": d["ms_pack_kernel"] / steps},
        "gpu_launches": d["kernel_launches"], "clocks": clocks,
        "e2e": {"value": T * steps / dt_enc, "unit": "frames/s", "h2d_bytes_per_step": T * FB,
                "d2h_bytes_per_step": sum(g["packet_bytes"] for g in gathered), "numa_node": numa_node,
                "note": "every rank: pinned host frames of its range -> host packets (submit_host / collect_async), wall clock "
                        "over all steps, max over ranks"},
        "decode": dec_line,
    }
