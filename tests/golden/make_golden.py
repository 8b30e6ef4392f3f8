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

def fate_goldens(reference="/root/reference"):
    """tests/golden/fate_vsynth.json: the md5 / size lines of the reference's own FATE goldens for this path
    (tests/ref/vsynth/vsynth{1,2,3}-ffv1{,-v0,-v3-yuv420p,-v3-yuv422p10,-v3-yuv444p16,-v3-bgr0}); test vectors only, used by tests/test_gpu_dropin_avcodec.py"""
    import json, os
    out = {}
    for clip in ("vsynth1", "vsynth2", "vsynth3"):
        for name in ("ffv1", "ffv1-v0", "ffv1-v3-yuv420p", "ffv1-v3-yuv422p10", "ffv1-v3-yuv444p16", "ffv1-v3-bgr0"):
            lines = open(os.path.join(reference, "tests/ref/vsynth/%s-%s" % (clip, name))).read().split("\n")
            out["%s-%s" % (clip, name)] = {"avi_md5": lines[0].split()[0], "avi_size": int(lines[1].split()[0]),
                                           "decoded_md5": lines[2].split()[0]}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fate_vsynth.json"), "w"), indent=1, sort_keys=True)
