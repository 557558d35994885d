"""Seeded synthetic inputs of the benchmark shapes.  numpy only.

There is no dataset in the repository or on the GPU box, so every test and bench line uses
these generators; ``seed`` is the frame index unless stated otherwise.
"""
import numpy as np

KITTI_SHAPE = (376, 1241)   # rows, cols  (Examples/Stereo/KITTI00-02.yaml)
TUM_SHAPE = (480, 640)      # rows, cols  (Examples/RGB-D/TUM1.yaml)
KITTI_FX = 718.856
KITTI_BF = 386.1448
TUM_BF = 40.0


def _gauss_blur_f32(img, sigma):
    r = max(1, int(3 * sigma + 0.5))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def blocky_canvas(shape, seed):
    """Noise-free float canvas: grey background with w*h/400 random grey rectangles, blurred."""
    h, w = shape
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128.0, np.float32)
    n = (w * h) // 400
    xs = rng.integers(0, w, n)
    ys = rng.integers(0, h, n)
    ws = rng.integers(4, 60, n)
    hs = rng.integers(4, 60, n)
    gs = rng.integers(0, 256, n)
    for x, y, ww, hh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + hh, x:x + ww] = g
    return _gauss_blur_f32(img, 0.8)


def _finish(canvas, rng):
    out = canvas + rng.normal(0.0, 2.0, canvas.shape).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def blocky_image(shape, seed):
    """The default generator: ~2.6k level-0 FAST candidates at KITTI shape."""
    return _finish(blocky_canvas(shape, seed), np.random.default_rng(seed + 1_000_003))


def noise_image(shape, seed):
    """White-noise stress case (~42k level-0 candidates at KITTI shape)."""
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def flat_image(shape, value=128):
    return np.full(shape, value, np.uint8)


def stereo_pair(shape, seed):
    """Left = blocky image; right = the same canvas shifted by a per-row-band disparity in
    [5, 80] px plus independent noise."""
    h, w = shape
    canvas = blocky_canvas(shape, seed)
    rng = np.random.default_rng(seed + 2_000_003)
    left = _finish(canvas, rng)
    band = 47
    nb = (h + band - 1) // band
    disp = rng.integers(5, 81, nb)
    right_c = np.empty_like(canvas)
    for b in range(nb):
        d = int(disp[b])
        rows = slice(b * band, min(h, (b + 1) * band))
        # a point at x in the left eye appears at x - d in the right eye
        right_c[rows, :w - d] = canvas[rows, d:]
        right_c[rows, w - d:] = canvas[rows, w - 1:w]
    right = _finish(right_c, rng)
    return left, right


def random_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def flip_bits(desc, nflip, rng):
    """Copy of ``desc`` (n x 32 u8) with nflip[i] random bits flipped in row i."""
    out = desc.copy()
    for i in range(out.shape[0]):
        k = int(nflip[i])
        if k:
            pos = rng.choice(256, size=k, replace=False)
            np.bitwise_xor.at(out[i], pos >> 3, (1 << (pos & 7)).astype(np.uint8))
    return out
