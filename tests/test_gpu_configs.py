"""BASELINE.json configs 3, 4 and 5 at their named sizes against SHA-256 manifests of the UNMODIFIED reference
(tests/golden/config{3,4,5}_manifest.json, made in the build container by tools/make_golden_configs.py: hours of CPU,
so the GPU box only hashes).  Bar: identical stream bytes and identical reconstruction for every picture.

  config 3: synthetic 768x512, qpd6=2                      -- pictures 0..255 (the whole manifest) in one batch call
  config 4: synthetic 3840x2160 -> 3840x2176, qpd6 0 and 4 -- 2 pictures per qpd6 by default, all 8 with HEVCE_SLOW=1
  config 5: one 15991x11993 picture, qpd6=2                -- (a) drop-in limit: top-left 8192x8192 (HEVCe.c:1581-1582),
                                                              (b) raised limit: padded to 16000x12000   [HEVCE_SLOW=1]
"""
import json
import os
import time

import numpy as np
import pytest

import refutil as R
import workloads as WL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SLOW = os.environ.get("HEVCE_SLOW") == "1"


@pytest.fixture(scope="module")
def H():
    import hevce_b200
    hevce_b200.set_variant(None)
    return hevce_b200


def manifest(name):
    p = os.path.join(GOLD, name + "_manifest.json")
    if not os.path.exists(p):
        pytest.skip(f"{name} manifest not generated")
    return json.load(open(p))["pictures"]


def check(entry, img, stream, rcon, what):
    assert R.sha(img.tobytes()) == entry["in_sha256"], f"{what}: the generator no longer produces the manifest's input"
    if list(rcon.shape) == entry["rcon_shape"] and "rcon_band_sha" in entry and R.sha(rcon.tobytes()) != entry["rcon_sha256"]:
        bands = [R.sha(rcon[y:y + 32].tobytes())[:12] for y in range(0, rcon.shape[0], 32)]
        first = next(i for i, (a, b) in enumerate(zip(bands, entry["rcon_band_sha"])) if a != b)
        pytest.fail(f"{what}: reconstruction differs from the reference from CTU row {first} on")
    assert list(rcon.shape) == entry["rcon_shape"], what
    assert len(stream) == entry["len"] and R.sha(stream) == entry["stream_sha256"], what
    assert R.sha(rcon.tobytes()) == entry["rcon_sha256"], what


def test_config3_against_reference_manifest(H):
    man = manifest("config3")
    n = len(man)                                   # all 256 pictures of the manifest: under a second of GPU time
    K = WL.kodak_landscape()
    imgs = [WL.config3_image(i, K) for i in range(n)]
    streams, rcons = H.HEVCImageEncoderBatch(imgs, 2)
    for i in range(n):
        check(man[f"{i:04d}"], imgs[i], streams[i], rcons[i], f"config 3 picture {i}")


@pytest.mark.parametrize("q", [0, 4])
def test_config4_against_reference_manifest(H, q):
    man = manifest("config4")
    ids = sorted(man, key=int)[: (8 if SLOW else 2)]
    K = WL.kodak_landscape()
    imgs = [WL.config4_image(int(i), K) for i in ids]
    t = time.time()
    streams, rcons = H.HEVCImageEncoderBatch(imgs, q)
    print(f"config 4, qpd6={q}: {len(ids)} pictures in {time.time() - t:.1f} s")
    for i, im, s, r in zip(ids, imgs, streams, rcons):
        check(man[i][f"q{q}"], im, s, r, f"config 4 picture {i} qpd6={q}")


@pytest.mark.slow
def test_config5_drop_in_crop(H):
    man = manifest("config5")
    if "crop" not in man:
        pytest.skip("config 5 (crop) not in the manifest")
    img = WL.config5_image()
    t = time.time()
    s, r = H.HEVCImageEncoder(img, 2)
    print(f"config 5 drop-in (8192x8192 of {img.shape[1]}x{img.shape[0]}): {time.time() - t:.1f} s, {len(s)} bytes")
    check(man["crop"], img, s, r, "config 5 (8192 crop)")


@pytest.mark.slow
def test_config5_raised_limit(H):
    man = manifest("config5")
    if "xl" not in man:
        pytest.skip("config 5 (raised limit) not in the manifest")
    img = WL.config5_image()
    old = H.set_max_dim(16384)
    try:
        t = time.time()
        s, r = H.HEVCImageEncoder(img, 2)
        print(f"config 5 raised limit ({img.shape[1]}x{img.shape[0]} -> {r.shape[1]}x{r.shape[0]}): {time.time() - t:.1f} s, {len(s)} bytes")
    finally:
        H.set_max_dim(old)
    check(man["xl"], img, s, r, "config 5 (16000x12000)")
