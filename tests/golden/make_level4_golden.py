#!/usr/bin/env python3
"""Generates tests/golden/level4.json from the UNMODIFIED reference build (oracle/_ref/libffv1ref.so) run with
-level 4 -strict experimental: extradata and (size, MD5, key flag) of every packet; for the cases whose reference
output cannot be reproduced (planar layouts, see tests/level4_cases.py) the reference's packets themselves
(zlib + base64), as input for the decoders.  Run in the build container: python tests/golden/make_level4_golden.py"""
import base64, hashlib, json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from level4_cases import LEVEL4_CASES, make_frames
from oracle import ffv1_ref as R

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

out = {}
for case in LEVEL4_CASES:
    cid, w, h, fmt, opts, n, exact = case
    frames = make_frames(case)
    e = R.Encoder(w, h, fmt, strict_experimental=1, **opts)
    pk = [e.encode(f) for f in frames]
    out[cid] = {"input_md5": md5(b"".join(f.tobytes() for f in frames)), "extradata": e.extradata.hex(),
                "packets": [[len(p), md5(p), int(k)] for p, k in pk]}
    if not exact:
        out[cid]["packets_z"] = [base64.b64encode(zlib.compress(p, 9)).decode() for p, k in pk[:2]]      # a keyframe and a non-keyframe
    print(cid, [len(p) for p, k in pk])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "level4.json"), "w"), indent=0, sort_keys=True)
