"""GOP-aligned frame-range partition of an encode/decode job across GPUs.

Model state only carries over inside a GOP (the reference resets it on keyframes, ffv1enc.c:1171-1172, 1299-1307), and
no pixel of another frame is ever read, so frame ranges that start on a keyframe are independent units: each rank gets
whole GOPs, encodes them with `first_picture_number` = the global index of its first frame, and the packets are put
back in pts order.  No collective is needed on the data path."""

def gop_aligned_ranges(nframes, gop_size, world_size):
    """[(first_frame, frame_count)] per rank; GOP counts differ by at most one, earlier ranks get the larger share."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    g = gop_size if gop_size > 0 else 1            # gop_size 0: every frame is a keyframe (ffv1enc.c:1299)
    ngops = (nframes + g - 1) // g
    base, extra = divmod(ngops, world_size)
    out, first_gop = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        start = min(first_gop * g, nframes)
        end = min((first_gop + n) * g, nframes)
        out.append((start, end - start))
        first_gop += n
    return out

def reinterleave(per_rank_packets):
    """per_rank_packets[r] = packets of rank r's range in coding order -> one list in pts order"""
    out = []
    for pk in per_rank_packets:
        out.extend(pk)
    return out
