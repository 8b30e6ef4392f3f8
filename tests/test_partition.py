"""Host-side multi-GPU logic on CPU: GOP-aligned frame ranges and the world-size-2 path over gloo.
Each rank encodes only its own range (here with the oracle standing in for the GPU codec -- the partition logic is what
is under test); the re-interleaved packets must equal a single-stream encode, i.e. nothing has to be exchanged."""
import hashlib, os, socket, sys
import numpy as np, pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
from ffv1_b200.partition import gop_aligned_ranges, reinterleave

def test_survey_example_split():
    # cfg5: 2400 frames, GOP 16 = 150 GOPs -> 19,19,19,19,19,19,18,18 GOPs on 8 GPUs
    r = gop_aligned_ranges(2400, 16, 8)
    assert [n // 16 for _, n in r] == [19, 19, 19, 19, 19, 19, 18, 18]
    assert [n // 16 for _, n in gop_aligned_ranges(2400, 16, 4)] == [38, 38, 37, 37]
    assert [n // 16 for _, n in gop_aligned_ranges(2400, 16, 2)] == [75, 75]
    assert all(s % 16 == 0 for s, _ in r) and sum(n for _, n in r) == 2400 and r[0][0] == 0
    for (s0, n0), (s1, _) in zip(r, r[1:]):
        assert s0 + n0 == s1

@pytest.mark.parametrize("nframes,gop,world", [(1, 16, 4), (17, 16, 2), (33, 16, 8), (10, 1, 3), (10, 0, 3), (100, 12, 8)])
def test_ragged_jobs(nframes, gop, world):
    r = gop_aligned_ranges(nframes, gop, world)
    assert len(r) == world and sum(n for _, n in r) == nframes
    g = gop if gop > 0 else 1
    pos = 0
    for s, n in r:
        assert s == pos and (s % g == 0 or n == 0)
        pos += n

def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import ffv1_oracle as O, synth
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    w, h, fmt, gop, nframes = 96, 64, "yuv420p", 4, 22
    opts = dict(level=3, coder=1, context=0, slices=4)
    gen = synth.Noisy(w, h, fmt, 99)
    frames = [gen.next() for _ in range(nframes)]          # every rank can regenerate the clip; it only codes its range
    start, count = gop_aligned_ranges(nframes, gop, world)[rank]
    enc = O.Encoder(w, h, fmt, gop=gop, **opts)            # a fresh encoder at a GOP boundary == continuing picture_number
    mine = [hashlib.md5(enc.encode(f)[0]).hexdigest() for f in frames[start:start + count]]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                 # control plane only (what bench.py does with its timings)
    if rank == 0:
        single = O.Encoder(w, h, fmt, gop=gop, **opts)
        expect = [hashlib.md5(single.encode(f)[0]).hexdigest() for f in frames]
        q.put(reinterleave(gathered) == expect)
    dist.barrier()
    dist.destroy_process_group()

def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
