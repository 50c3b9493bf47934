"""Deterministic synthetic test images: three content classes mixed 1:1:1 so that flat,
textured and noisy macroblocks (skip / I16 / I4 paths, all segments) all occur.  No files, no network."""
import numpy as np


def synth_image(w, h, index, kind=None):
    """Opaque RGBA uint8 [h][w][4]; kind 0 gradient (richTestImage, root encode_test.go:1496), 1 gradient +
    band-limited noise + hard-edged rectangles, 2 noisyImage (race_test.go:78)."""
    kind = index % 3 if kind is None else kind
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    rng = np.random.RandomState(0xC0FFEE + index)
    if kind == 0:
        r = xx * 255 // max(w, 1)
        g = yy * 255 // max(h, 1)
        b = (xx + yy) * 255 // max(w + h, 1)
    elif kind == 1:
        base = rng.randint(-24, 25, size=((h + 7) // 8 + 1, (w + 7) // 8 + 1, 3))
        noise = np.kron(base, np.ones((8, 8, 1), np.int64))[:h, :w]
        r = xx * 255 // max(w, 1) + noise[..., 0]
        g = yy * 255 // max(h, 1) + noise[..., 1]
        b = 128 + noise[..., 2]
        for _ in range(12):
            x0, y0 = rng.randint(0, w), rng.randint(0, h)
            x1, y1 = min(w, x0 + rng.randint(4, max(5, w // 4))), min(h, y0 + rng.randint(4, max(5, h // 4)))
            col = rng.randint(0, 256, 3)
            r[y0:y1, x0:x1], g[y0:y1, x0:x1], b[y0:y1, x0:x1] = col
    else:
        r = (xx * 7 + yy * 13) % 256 + rng.randint(-16, 17, size=(h, w))
        g = (xx * 3 + yy * 5) % 256 + rng.randint(-16, 17, size=(h, w))
        b = (xx ^ yy) % 256
    img = np.stack([r, g, b, np.full_like(r, 255)], axis=-1)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_batch(n, w, h, distinct=24, first_index=0):
    """n images; `distinct` different ones generated, then repeated (generation is numpy-bound, not the workload)."""
    d = min(n, distinct)
    base = np.stack([synth_image(w, h, first_index + i) for i in range(d)])
    reps = (n + d - 1) // d
    return np.concatenate([base] * reps)[:n]


def synth_formula(w, h, kind):
    """RNG-free pictures (integer formulas only) that a Go program reproduces bit for bit (go/cmd/refdump): kind 0 = the
    gradient of richTestImage (root encode_test.go:1496), 1 = noisyImage without its random term (race_test.go:78), 2 = both
    in alternating 32x32 tiles (hard edges).  Opaque RGBA uint8 [h][w][4]."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    grad = (xx * 255 // w, yy * 255 // h, (xx + yy) * 255 // (w + h))
    noisy = ((xx * 7 + yy * 13) % 256, (xx * 3 + yy * 5) % 256, (xx ^ yy) % 256)
    if kind == 0:
        r, g, b = grad
    elif kind == 1:
        r, g, b = noisy
    else:
        sel = (((xx >> 5) + (yy >> 5)) & 1) == 0
        r, g, b = [np.where(sel, a, c) for a, c in zip(grad, noisy)]
    return np.stack([r, g, b, np.full_like(r, 255)], axis=-1).astype(np.uint8)
