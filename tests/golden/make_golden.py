#!/usr/bin/env python3
"""Generates tests/golden/ref_packets.json from the UNMODIFIED reference build (oracle/_ref/libffv1ref.so).
Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
For every case of tests/cases.py (+ the 1080p S2 'noisy1080' clip of SURVEY 8(d)) it records the MD5 of the
synthetic input, the extradata MD5 and (size, MD5, key flag) of every packet the reference encoder produced."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES, make_frames
from oracle import ffv1_ref as R, synth

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

def run(w, h, fmt, opts, frames):
    e = R.Encoder(w, h, fmt, **opts)
    rec = {"w": w, "h": h, "pix_fmt": fmt, "opts": opts, "extradata_md5": md5(e.extradata),
           "extradata_size": len(e.extradata), "input_md5": md5(b"".join(f.tobytes() for f in frames)), "packets": []}
    for f in frames:
        pkt, key = e.encode(f)
        rec["packets"].append([len(pkt), md5(pkt), int(key)])
    return rec

out = {}
for case in CASES:
    cid, w, h, fmt, opts, kind, n = case
    out[cid] = run(w, h, fmt, opts, make_frames(case))
    out[cid]["kind"] = kind
g = synth.Noisy(1920, 1080, "yuv420p", 1234)
out["s2_noisy1080_c2"] = run(1920, 1080, "yuv420p", dict(gop=16, level=3, coder=1, context=0, slices=24), [g.next() for _ in range(32)])
out["s2_noisy1080_c2"]["kind"] = "s2"
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_packets.json"), "w"), indent=0, sort_keys=True)
print("wrote", len(out), "cases")
