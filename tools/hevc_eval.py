#!/usr/bin/env python
"""Quality study on the batch API: what the reference's HEVCeval.py measures
(HEVCeval.py:119-245) -- bits per pixel of HEVC against JPEG / JPEG2000 / WebP at equal SSIM -- with every picture of
the input directory encoded in ONE HEVCImageEncoderBatch call instead of one CLI process (and three 2-second sleeps)
per picture.

    python tools/hevc_eval.py <input-dir> <output-dir> [<qpd6>]          (default qpd6 = 3, as the reference)

Same conventions as the reference script: pictures are converted to 8-bit monochrome, padded to multiples of 32 by
edge replication before anything is measured, `<name>.h265` and the best-matching `<name>.jpg/.j2k/.webp` are written
to the output directory, the per-format quality parameter is found by bisection on SSIM (data_range = 256).
SSIM is computed here (7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03 -- the defaults of the
scikit-image function the reference calls), so scikit-image is not needed.  Needs a CUDA device: there is no CPU
encoder in this repository.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hevc-image-encoder-lite_b200"))

COMPARISONS = [("JPEG", ".jpg", 1, 101), ("JPEG2000", ".j2k", 25, 75), ("WEBP", ".webp", 1, 101)]   # HEVCeval.py:100-105


def image_pad(img, pad=32):
    """Pad to multiples of `pad` by edge replication (HEVCeval.py:20-41)."""
    h, w = img.shape
    return np.pad(img, ((0, (-h) % pad), (0, (-w) % pad)), mode="edge")


def read_monochrome(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("L"))


def ssim(a, b, data_range=256.0, win=7):
    """Mean structural similarity with scikit-image's default settings (uniform 7x7 window, sample covariance)."""
    from scipy.ndimage import uniform_filter
    a, b = a.astype(np.float64), b.astype(np.float64)
    n = win * win
    cov_norm = n / (n - 1.0)
    ua, ub = uniform_filter(a, win), uniform_filter(b, win)
    va = cov_norm * (uniform_filter(a * a, win) - ua * ua)
    vb = cov_norm * (uniform_filter(b * b, win) - ub * ub)
    vab = cov_norm * (uniform_filter(a * b, win) - ua * ub)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ua * ub + c1) * (2 * vab + c2)) / ((ua * ua + ub * ub + c1) * (va + vb + c2))
    p = (win - 1) // 2
    return float(s[p:s.shape[0] - p, p:s.shape[1] - p].mean())


def save_as(img, path, quality):
    """HEVCeval.py:86-95."""
    from PIL import Image
    im = Image.fromarray(img)
    if path.endswith(".j2k"):
        im.save(path, optimize=True, quality_mode="dB", quality_layers=[quality])
    else:
        im.save(path, optimize=True, quality=quality)


def match_ssim(img, path, target, lo, hi):
    """Bisection on the quality parameter for the SSIM nearest to `target` (HEVCeval.py:199-224)."""
    tried = []
    while hi - lo > 1:
        q = (hi + lo) // 2
        save_as(img, path, q)
        s = ssim(img, read_monochrome(path))
        tried.append((abs(s - target), s, os.path.getsize(path), q))
        if s < target:
            lo = q
        else:
            hi = q
    tried.sort(key=lambda t: t[0])
    _, s, size, q = tried[0]
    save_as(img, path, q)
    return s, size, q


def evaluate(in_dir, out_dir, qpd6=3, formats=COMPARISONS, log=print):
    import hevce_b200 as H
    os.makedirs(out_dir, exist_ok=True)
    names, imgs = [], []
    for fname in sorted(os.listdir(in_dir)):
        try:
            imgs.append(image_pad(read_monochrome(os.path.join(in_dir, fname))))
            names.append(fname)
        except Exception:
            continue
    streams, rcons = H.HEVCImageEncoderBatch(imgs, qpd6)            # the whole directory in one call
    rows, bpp = [], {"HEVC": []}
    for fname, img, stream, rcon in zip(names, imgs, streams, rcons):
        h, w = img.shape
        base = os.path.join(out_dir, os.path.splitext(fname)[0])
        with open(base + ".h265", "wb") as f:
            f.write(stream)
        hevc_ssim, hevc_bpp = ssim(img, rcon), 8.0 * len(stream) / (w * h)
        bpp["HEVC"].append(hevc_bpp)
        log("\n%s    width=%d    height=%d" % (os.path.join(in_dir, fname), w, h))
        log("  HEVC     : ssim=%.5f    bpp=%.3f" % (hevc_ssim, hevc_bpp))
        row = {"name": fname, "HEVC": (hevc_ssim, hevc_bpp)}
        for name, suffix, lo, hi in formats:
            s, size, q = match_ssim(img, base + suffix, hevc_ssim, lo, hi)
            b = 8.0 * size / (w * h)
            bpp.setdefault(name, []).append(b)
            row[name] = (s, b, q)
            log("  %-8s : ssim=%.5f    bpp=%.3f    qparam=%d    size/HEVCsize=%f" % (name, s, b, q, size / len(stream)))
        rows.append(row)
        log("bpp mean : " + "   ".join("%s:%.5f" % (k, np.mean(v)) for k, v in bpp.items()))
    return rows, {k: float(np.mean(v)) for k, v in bpp.items() if v}


if __name__ == "__main__":
    try:
        in_dir, out_dir = sys.argv[1:3]
        assert in_dir != out_dir
    except Exception:
        print("\n    Usage:\n        python  %s  <input-dir>  <output-dir>  [<qpd6>]\n" % sys.argv[0])
        sys.exit(-1)
    q = 3
    try:
        q = int(sys.argv[3])
    except Exception:
        pass
    print("\n|-arguments --------------------------------------")
    print("|   input  dir     = %s" % in_dir)
    print("|   output dir     = %s" % out_dir)
    print("|   Qp%%6           = %d        (Qp = %d)" % (q, q * 6 + 4))
    print("|-------------------------------------------------\n")
    _, means = evaluate(in_dir, out_dir, q)
    print("\n\nbpp means ---------------------------------")
    for k, v in means.items():
        print("%-8s %.5f   (%+.1f %% vs HEVC)" % (k, v, 100.0 * (v / means["HEVC"] - 1.0)))
