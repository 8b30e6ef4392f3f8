"""Pixel-format geometry shared by the oracle/reference test wrappers (test infrastructure only).

name -> (bytes_per_sample, [(h_shift, v_shift) per stored plane], packed_pixel_bytes)
Matches the formats the reference encoder accepts (ffv1enc.c:1425-1439)."""

def describe(name):
    bps = 2 if name.endswith("le") else 1
    if name in ("bgr0", "bgra"):
        return 1, [(0, 0)], 4
    if name == "ya8":
        return 1, [(0, 0)], 2
    if name.startswith("gray"):
        return bps, [(0, 0)], 1
    if name.startswith("gbrp"):
        return bps, [(0, 0)] * 3, 1
    sub = {"420": (1, 1), "422": (1, 0), "444": (0, 0), "440": (0, 1), "411": (2, 0), "410": (2, 2)}
    for k, v in sub.items():
        if k in name:
            planes = [(0, 0), v, v]
            if name.startswith("yuva"):
                planes.append((0, 0))
            return bps, planes, 1
    raise ValueError(name)

def plane_shapes(name, w, h):
    """[(rows, row_bytes)] for tightly packed planes"""
    bps, planes, ppb = describe(name)
    out = []
    for hs, vs in planes:
        cw = -((-w) >> hs)
        ch = -((-h) >> vs)
        out.append((ch, cw * bps * ppb))
    return out

def frame_bytes(name, w, h):
    return sum(r * b for r, b in plane_shapes(name, w, h))
