#!/usr/bin/env python3
"""Generates tests/golden/twopass.json from the UNMODIFIED reference build (oracle/_ref/libffv1ref.so): for every case of
tests/twopass_cases.py the first pass's stats_out text (zlib + base64), and the second pass's extradata and packets
(size, MD5, key flag).  Run in the build container:  python tests/golden/make_twopass_golden.py"""
import base64, hashlib, json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from twopass_cases import TWOPASS_CASES, make_frames
from oracle import ffv1_ref as R

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

out = {}
for case in TWOPASS_CASES:
    cid, w, h, fmt, opts, n = case
    frames = make_frames(case)
    e1 = R.Encoder(w, h, fmt, two_pass=1, **opts)
    p1 = [e1.encode(f) for f in frames]
    stats = e1.stats_out()
    e2 = R.Encoder(w, h, fmt, two_pass=2, stats_in=stats, **opts)
    p2 = [e2.encode(f) for f in frames]
    out[cid] = {"input_md5": md5(b"".join(f.tobytes() for f in frames)),
                "pass1_extradata": e1.extradata.hex(), "pass1_packets": [[len(p), md5(p), int(k)] for p, k in p1],
                "stats_z": base64.b64encode(zlib.compress(stats.encode(), 9)).decode(), "stats_md5": md5(stats.encode()),
                "pass2_extradata": e2.extradata.hex(), "pass2_packets": [[len(p), md5(p), int(k)] for p, k in p2]}
    print(cid, len(stats), len(out[cid]["stats_z"]), sum(x[0] for x in out[cid]["pass1_packets"]), sum(x[0] for x in out[cid]["pass2_packets"]))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "twopass.json"), "w"), indent=0, sort_keys=True)
