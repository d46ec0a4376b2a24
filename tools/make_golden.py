#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the build container only).

Inputs : /root/reference/testimage/NN.pgm (24 Kodak grayscale images), /root/reference/testimage_out/NN.h265
         (the reference's own shipped goldens, made at qpd6=4) and oracle/_ref/libhevce_ref.so (the
         unmodified reference compiled by `make -C oracle`).
Outputs: tests/golden/kodak_gray.npz        the 24 input images (lossless, compressed) -- the GPU box has no
                                            /root/reference, so the inputs have to travel with the repo
         tests/golden/kodak_manifest.json   len + SHA-256 of every stream / reconstruction, 24 images x qpd6 0..4,
                                            plus SHA-256 of the shipped testimage_out/*.h265 files
         tests/golden/small_cases.npz/.json small inputs with their full expected streams and reconstructions

usage: python tools/make_golden.py [--jobs 8] [--skip-kodak]
"""
import argparse
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as R  # noqa: E402

REFDIR = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def small_case_inputs(k01):
    """Deterministic small inputs (name -> 2-D uint8 array). k01 = Kodak image 01 (512x768)."""
    rng = np.random.default_rng(0)
    cases = {
        "k01_1x1": k01[0:1, 0:1],
        "k01_32x32": k01[0:32, 0:32],
        "k01_45x70": k01[0:45, 0:70],
        "k01_64x64": k01[0:64, 0:64],
        "k01_64x128": k01[100:164, 200:328],
        "k01_33x97": k01[300:333, 400:497],
        "k01_96x64": k01[200:296, 640:704],
        "flat128_64": np.full((64, 64), 128, np.uint8),
        "flat0_64": np.zeros((64, 64), np.uint8),
        "flat255_64": np.full((64, 64), 255, np.uint8),
        "noise_64": rng.integers(0, 256, (64, 64)).astype(np.uint8),
        "binary_64": (rng.integers(0, 2, (64, 64)) * 255).astype(np.uint8),
        "checker1_64": ((np.indices((64, 64)).sum(0) & 1) * 255).astype(np.uint8),
        "checker2_64": (((np.indices((64, 64)) // 2).sum(0) & 1) * 255).astype(np.uint8),
        "vstripes_64": np.tile((np.arange(64) & 1) * 255, (64, 1)).astype(np.uint8),
        "gauss_64": np.clip(np.round(rng.normal(128, 20, (64, 64))), 0, 255).astype(np.uint8),
        "ramp_40x100": (np.add.outer(np.arange(40) * 3, np.arange(100) * 2) % 256).astype(np.uint8),
    }
    return cases


def _encode(args):
    name, img, q = args
    s, r = R.ref_encode(img, q)
    return name, q, s, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--skip-kodak", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    R.build_oracles() if not os.path.exists(R.REF_SO) else None

    kodak = {f"{i:02d}": R.read_pgm(f"{REFDIR}/testimage/{i:02d}.pgm") for i in range(1, 25)}
    np.savez_compressed(os.path.join(GOLD, "kodak_gray.npz"), **{"k" + k: v for k, v in kodak.items()})

    # ---- small cases: full streams + recon
    cases = small_case_inputs(kodak["01"])
    jobs = [(n, img, q) for n, img in cases.items() for q in range(5)]
    out, meta = {}, {}
    with ProcessPoolExecutor(a.jobs) as ex:
        for name, q, s, r in ex.map(_encode, jobs):
            out[f"{name}/in"] = cases[name]
            out[f"{name}/q{q}/stream"] = np.frombuffer(s, np.uint8)
            out[f"{name}/q{q}/rcon"] = r
            meta.setdefault(name, {"shape": list(cases[name].shape), "q": {}})["q"][str(q)] = {
                "len": len(s), "stream_sha256": R.sha(s), "rcon_sha256": R.sha(r.tobytes())}
    np.savez_compressed(os.path.join(GOLD, "small_cases.npz"), **out)
    json.dump(meta, open(os.path.join(GOLD, "small_cases.json"), "w"), indent=1, sort_keys=True)
    print("small cases:", len(cases), "inputs x 5 qpd6")

    if a.skip_kodak:
        return
    # ---- Kodak 24 x 5
    man = {}
    for k, img in kodak.items():
        shipped = open(f"{REFDIR}/testimage_out/{k}.h265", "rb").read()
        man[k] = {"shape": list(img.shape), "shipped_q4_len": len(shipped), "shipped_q4_sha256": R.sha(shipped), "q": {}}
    jobs = [(k, img, q) for q in range(5) for k, img in kodak.items()]
    with ProcessPoolExecutor(a.jobs) as ex:
        for name, q, s, r in ex.map(_encode, jobs):
            h, w = r.shape
            pgm = b"P5\n%d %d\n255\n" % (w, h) + r.tobytes()
            man[name]["q"][str(q)] = {"len": len(s), "stream_sha256": R.sha(s), "rcon_sha256": R.sha(r.tobytes()),
                                      "rcon_pgm_sha256": R.sha(pgm)}
            if q == 4:
                assert R.sha(s) == man[name]["shipped_q4_sha256"], f"oracle/_ref disagrees with shipped golden {name}"
            print(name, q, len(s), flush=True)
    json.dump(man, open(os.path.join(GOLD, "kodak_manifest.json"), "w"), indent=1, sort_keys=True)
    print("kodak manifest written; all 24 shipped qpd6=4 goldens reproduced")


if __name__ == "__main__":
    main()
