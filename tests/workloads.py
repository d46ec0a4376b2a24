"""Deterministic synthetic workloads of the BASELINE.json configs, built from the committed
Kodak fixtures so they exist on the GPU box (which has no /root/reference)."""
import numpy as np

import golden_util as G

KODAK_H, KODAK_W = 512, 768


def kodak_landscape():
    """K[1..24] as 512x768 arrays (the six portrait pictures transposed)."""
    imgs, _ = G.kodak()
    out = []
    for j in range(1, 25):
        a = imgs[f"k{j:02d}"]
        out.append(np.ascontiguousarray(a.T) if a.shape == (KODAK_W, KODAK_H) else a)
    return out


def config3_image(i, K=None):
    """Image i of config 3: K[(i mod 24)+1] circularly shifted by (dy,dx) from default_rng(1234+i) (dy in [0,512),
    dx in [0,768), drawn in that order) plus i.i.d. noise integers(-2,3), clipped to [0,255]. 768x512 (w x h)."""
    K = K or kodak_landscape()
    rng = np.random.default_rng(1234 + i)
    dy = int(rng.integers(0, KODAK_H))
    dx = int(rng.integers(0, KODAK_W))
    base = np.roll(K[i % 24], (dy, dx), axis=(0, 1)).astype(np.int16)
    noise = rng.integers(-2, 3, size=base.shape)
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def config3_batch(first, count):
    K = kodak_landscape()
    return [config3_image(i, K) for i in range(first, first + count)]


def config4_image(i, K=None, h=2160, w=3840, seed=4000):
    """Image i of config 4: 5x5 mosaic of Kodak tiles (tile choice integers(1,25), per-tile h/v flips from
    default_rng(4000+i)), cropped to h x w."""
    K = K or kodak_landscape()
    rng = np.random.default_rng(seed + i)
    ty, tx = -(-h // KODAK_H), -(-w // KODAK_W)
    rows = []
    for _ in range(ty):
        row = []
        for _ in range(tx):
            t = K[int(rng.integers(1, 25)) - 1]
            if rng.integers(0, 2):
                t = t[:, ::-1]
            if rng.integers(0, 2):
                t = t[::-1, :]
            row.append(t)
        rows.append(np.concatenate(row, axis=1))
    return np.ascontiguousarray(np.concatenate(rows, axis=0)[:h, :w])


CONFIG5_H, CONFIG5_W = 11993, 15991


def config5_image(K=None, h=CONFIG5_H, w=CONFIG5_W):
    """The picture of config 5: the config-4 construction as a 21x24 mosaic from default_rng(5000),
    cropped to 15991x11993 (w x h), which the encoder pads to 16000x12000 (raised limit) or crops to the top-left
    8192x8192 (drop-in limit, HEVCe.c:1581-1582)."""
    return config4_image(0, K, h=h, w=w, seed=5000)


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n independent pictures for `rank` of `world` (no collective on the data path)."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi
