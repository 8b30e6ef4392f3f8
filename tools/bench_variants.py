"""Encode throughput of the non-headline variants on the bench clip (256-frame batches, host path): Golomb-Rice with the
small context model and the range coder with the large context model.  usage: bench_variants.py"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import numpy as np, ffv1_b200
from oracle import synth
W,H,FMT=1920,1080,"yuv420p"
g=synth.Noisy(W,H,FMT,1234); base=[g.next() for _ in range(16)]
n=256; frames=[base[i%16] for i in range(n)]
for coder, ctx in ((0,0),(1,1)):
    enc=ffv1_b200.FFV1Encoder(W,H,FMT,g=16,level=3,coder=coder,context=ctx,slices=24,max_batch_frames=n)
    enc.encode_batch(frames)
    t0=time.perf_counter(); p=enc.encode_batch(frames); dt=time.perf_counter()-t0
    s=enc.stats(); st={k:getattr(s,k) for k,_ in s._fields_}
    print("coder",coder,"context",ctx,"fps %.1f"%(n/dt), {k:round(v,1) for k,v in st.items() if k.startswith("ms_")}, "bytes/frame", sum(len(x[0]) for x in p)//n)
    enc.close()
