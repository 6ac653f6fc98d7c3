"""Deterministic synthetic grayscale frames for parity tests and benchmarks (numpy only).

Recipe (restated without cv2 so it runs identically everywhere): a canvas of
(H+128)x(W+128) filled with 128 receives n = 2600*H*W/1e6 axis-aligned rectangles (position
uniform over the canvas, sides in [6,48), grey in [0,256)), is smoothed with the 5x5 binomial
kernel [1 4 6 4 1]^2/256 (round-half-up), gets integer noise uniform in [-3,3] and is cropped at
offset 64+shift.  Frames of one sequence share `seed` (the canvas) and differ by `shift` and
`noise_seed`, so consecutive frames really do match.  `lowtex=True` compresses contrast to
128+-6 inside seeded rectangles covering ~25 % of the frame and flattens one band completely, so
that many FAST cells take the minThFAST fallback and some yield no corner at all.
"""
import numpy as np

MARGIN = 64


def _binomial5(img):
    k = np.array([1, 4, 6, 4, 1], dtype=np.int32)
    a = img.astype(np.int32)
    p = np.pad(a, ((0, 0), (2, 2)), mode="reflect")
    h = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(5))
    p = np.pad(h, ((2, 2), (0, 0)), mode="reflect")
    v = sum(k[i] * p[i:i + a.shape[0], :] for i in range(5))
    return ((v + 128) >> 8).astype(np.int32)


def canvas(H, W, seed, lowtex=False):
    rng = np.random.default_rng(seed)
    ch, cw = H + 2 * MARGIN, W + 2 * MARGIN
    c = np.full((ch, cw), 128, dtype=np.int32)
    n = int(2600 * H * W / 1e6)
    xs = rng.integers(0, cw, n); ys = rng.integers(0, ch, n)
    ws = rng.integers(6, 48, n); hs = rng.integers(6, 48, n)
    gs = rng.integers(0, 256, n)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        c[y:y + h, x:x + w] = g
    c = _binomial5(c)
    if lowtex:
        m = int(0.25 * ch * cw / (96 * 96)) + 1
        for _ in range(m):
            x = int(rng.integers(0, cw)); y = int(rng.integers(0, ch))
            w = int(rng.integers(64, 128)); h = int(rng.integers(64, 128))
            blk = c[y:y + h, x:x + w]
            c[y:y + h, x:x + w] = 128 + (blk - 128) * 6 // 128
        y0 = int(rng.integers(0, max(1, ch - int(0.15 * ch))))
        c[y0:y0 + int(0.15 * ch), :] = 117
    return c


def frame_from_canvas(c, H, W, shift=(0, 0), noise_seed=0, noise=True):
    sx, sy = shift
    assert -MARGIN <= sx <= MARGIN and -MARGIN <= sy <= MARGIN
    crop = c[MARGIN + sy:MARGIN + sy + H, MARGIN + sx:MARGIN + sx + W]
    if noise:
        rng = np.random.default_rng(noise_seed)
        crop = crop + rng.integers(-3, 4, size=(H, W))
        if (crop.max() <= 117 + 3) and (crop.min() >= 117 - 3):
            pass
    return np.clip(crop, 0, 255).astype(np.uint8)


def synth(H, W, seed, shift=(0, 0), noise_seed=None, lowtex=False):
    """One HxW uint8 frame."""
    c = canvas(H, W, seed, lowtex)
    return frame_from_canvas(c, H, W, shift, seed if noise_seed is None else noise_seed)


def sequence(H, W, seed, n, step=(2, 1), lowtex=False):
    """n frames of one scene, frame t shifted by t*step (clamped to the canvas margin)."""
    c = canvas(H, W, seed, lowtex)
    out = np.empty((n, H, W), dtype=np.uint8)
    for t in range(n):
        sx = int(np.clip(t * step[0], -MARGIN, MARGIN)); sy = int(np.clip(t * step[1], -MARGIN, MARGIN))
        out[t] = frame_from_canvas(c, H, W, (sx, sy), noise_seed=seed * 1000003 + t)
    return out


SHAPES = {  # BASELINE.json configs: name -> (H, W, nFeatures)
    "euroc": (480, 752, 1000),
    "kitti": (376, 1241, 2000),
    "tum": (480, 640, 1000),
    "4k": (2160, 3840, 8000),
}
