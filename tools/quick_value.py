#!/usr/bin/env python3
"""The `value` leg of bench.py alone (frames resident in HBM, CUDA events), for A/B runs of kernel variants:
prints frames/s and the per-stage kernel times of a step.  usage: quick_value.py [--batch N] [--steps K]
Environment variables of the library (FFV1B200_REPLAY_WINDOW, ...) select the variant."""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", str(a.batch), "--steps", str(a.steps), "--warmup", "1",
                      "--no-cpu-baseline", "--no-e2e", "--no-decode", "--no-avcodec"], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:]); sys.exit(1)
d = json.loads(line[-1])
print(round(d["value"]), {k: round(v, 2) for k, v in d["kernels_ms_per_step"].items()}, d.get("parity_checked"), d["clocks"])
