"""Deterministic synthetic inputs -- shared by tests and bench (generators only; no codec logic).
All return a contiguous uint8 array holding one tightly packed frame in the given pix_fmt."""
import numpy as np

class Noisy:
    """S2/S3/S4-style 'camera noise' clips: smooth ramps + gaussian noise, seeded; frames must be drawn in order."""
    def __init__(self, w, h, pix_fmt, seed):
        self.w, self.h, self.pix_fmt = w, h, pix_fmt
        self.rng = np.random.default_rng(seed)
        self.yy, self.xx = np.mgrid[0:h, 0:w]
        self.n = 0

    def next(self):
        n, rng, xx, yy, w, h, f = self.n, self.rng, self.xx, self.yy, self.w, self.h, self.pix_fmt
        self.n += 1
        base = (0.1 * xx + 0.07 * yy + 1.5 * n) % 256
        if f == "yuv420p":     # S2 "noisy1080" recipe, exactly as SURVEY 8(d)
            Y = np.clip(base + rng.normal(0, 2, (h, w)), 0, 255).astype(np.uint8)
            cx, cy = xx[::2, ::2], yy[::2, ::2]
            U = np.clip(128 + 20 * np.sin((cx + 3 * n) / 97) + rng.normal(0, 1.5, cx.shape), 0, 255).astype(np.uint8)
            V = np.clip(128 + 20 * np.cos((cy + 2 * n) / 71) + rng.normal(0, 1.5, cy.shape), 0, 255).astype(np.uint8)
            return np.concatenate([Y.ravel(), U.ravel(), V.ravel()])
        if f.startswith("gbrp"):
            bits = int(f[4:].rstrip("le"))
            mx = (1 << bits) - 1
            sc = (1 << bits) / 256.0
            G = np.clip(sc * base + rng.normal(0, sc / 2, (h, w)), 0, mx)
            B = np.clip(G + rng.normal(0, sc * 3 / 8, (h, w)), 0, mx)
            R = np.clip(G + rng.normal(0, sc * 3 / 8, (h, w)), 0, mx)
            return np.concatenate([p.astype("<u2").ravel() for p in (G, B, R)]).view(np.uint8)
        if f in ("bgr0", "bgra"):
            G = np.clip(base + rng.normal(0, 2, (h, w)), 0, 255)
            B = np.clip(G + rng.normal(0, 3, (h, w)), 0, 255)
            R = np.clip(G + rng.normal(0, 3, (h, w)), 0, 255)
            A = np.clip(200 + rng.normal(0, 2, (h, w)), 0, 255) if f == "bgra" else np.zeros((h, w))
            return np.stack([B, G, R, A], -1).astype(np.uint8).ravel()
        if f == "ya8":
            Y = np.clip(base + rng.normal(0, 2, (h, w)), 0, 255)
            A = np.clip(128 + 40 * np.sin(xx / 31.0) + rng.normal(0, 1, (h, w)), 0, 255)
            return np.stack([Y, A], -1).astype(np.uint8).ravel()
        # generic planar YUV(A)/gray, 8..16 bit
        from . import pixfmt
        bps, planes, _ = pixfmt.describe(f)
        bits = 8
        if bps == 2:
            digits = "".join(ch for ch in f.split("p")[-1] if ch.isdigit()) if "p" in f else "16"
            bits = int(digits) if digits else 16
            if f.startswith("gray"):
                bits = 16
        mx = (1 << bits) - 1
        sc = (1 << bits) / 256.0
        out = []
        for i, (hs, vs) in enumerate(planes):
            px, py = xx[::1 << vs, ::1 << hs], yy[::1 << vs, ::1 << hs]
            if i == 0:
                v = sc * ((0.1 * px + 0.07 * py + 1.5 * n) % 256) + rng.normal(0, 2 * sc, px.shape)
            elif i == 1:
                v = sc * (128 + 20 * np.sin((px + 3 * n) / 97)) + rng.normal(0, 1.5 * sc, px.shape)
            elif i == 2:
                v = sc * (128 + 20 * np.cos((py + 2 * n) / 71)) + rng.normal(0, 1.5 * sc, px.shape)
            else:
                v = sc * (200 + 30 * np.sin((px + py) / 53.0)) + rng.normal(0, sc, px.shape)
            v = np.clip(v, 0, mx)
            out.append(v.astype(np.uint8).ravel() if bps == 1 else v.astype("<u2").ravel().view(np.uint8))
        return np.concatenate(out)

def flat_bars(w, h, pix_fmt, n):
    """easy content (long zero runs, context 0 dominates): moving colour bars; exercises golomb run mode."""
    from . import pixfmt
    bps, planes, ppb = pixfmt.describe(pix_fmt)
    out = []
    for i, (hs, vs) in enumerate(planes):
        cw, ch = -((-w) >> hs), -((-h) >> vs)
        xs = (np.arange(cw * ppb) // ppb + 3 * n) // max(1, cw // 7)
        row = ((xs * (37 + 11 * i)) % 200 + 20)
        pl = np.tile(row, (ch, 1))
        pl[ch // 2:, :] = (pl[ch // 2:, :] + (np.arange(ch - ch // 2)[:, None] // 9) * 3) % 251
        if bps == 2:
            shift = 0
            out.append((pl.astype("<u2") << shift).ravel().view(np.uint8))
        else:
            out.append(pl.astype(np.uint8).ravel())
    return np.concatenate(out)

def random_frame(w, h, pix_fmt, seed, maxval=None):
    """uniform random samples: worst case for the coder (every escape / large-exponent branch)."""
    from . import pixfmt
    rng = np.random.default_rng(seed)
    bps, planes, ppb = pixfmt.describe(pix_fmt)
    out = []
    for hs, vs in planes:
        cw, ch = -((-w) >> hs), -((-h) >> vs)
        if bps == 1:
            out.append(rng.integers(0, 256, ch * cw * ppb, dtype=np.uint8))
        else:
            mv = maxval if maxval is not None else 65535
            out.append(rng.integers(0, mv + 1, ch * cw, dtype=np.uint16).astype("<u2").view(np.uint8))
    return np.concatenate(out)
