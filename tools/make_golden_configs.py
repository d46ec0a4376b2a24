#!/usr/bin/env python
"""SHA-256 manifests of the reference's output for BASELINE.json configs 3, 4 and 5 (build container only).

Every encode runs the UNMODIFIED reference (oracle/_ref/libhevce_ref.so, or libhevce_ref_xl.so = the same file with
its two size limits raised to 16384 for config 5b; both built by `make -C oracle`) on the deterministic pictures of
tests/workloads.py.  One result file per encode is kept under --work so an interrupted run resumes; `--merge` writes

  tests/golden/config3_manifest.json   pictures 0..N-1 of config 3 at qpd6=2
  tests/golden/config4_manifest.json   pictures 0..M-1 of config 4 at qpd6 0 and 4
  tests/golden/config5_manifest.json   "crop" (drop-in limit, 8192x8192) and "xl" (16000x12000) at qpd6=2

Each entry: stream length + SHA-256, reconstruction SHA-256, and for the large pictures the SHA-256 prefix of every
32-row band of the reconstruction (to localise a first mismatch).

usage: nice python tools/make_golden_configs.py --jobs 6 [--n3 256] [--n4 8]      (hours of CPU)
       python tools/make_golden_configs.py --merge
"""
import argparse
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as R  # noqa: E402
import workloads as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_XL = os.path.join(R.ORACLE_DIR, "_ref", "libhevce_ref_xl.so")


def _encode(lib_path, img, q, limit):
    lib = R._load(lib_path)
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    hp, wp = R.padded(h, limit), R.padded(w, limit)
    rcon = np.zeros((hp, wp), np.uint8)
    buf = np.zeros(256 + hp * wp * 2, np.uint8)
    ys, xs = ctypes.c_int(h), ctypes.c_int(w)
    u8p = ctypes.POINTER(ctypes.c_ubyte)
    n = lib.HEVCImageEncoder(buf.ctypes.data_as(u8p), img.ctypes.data_as(u8p), rcon.ctypes.data_as(u8p),
                             ctypes.byref(ys), ctypes.byref(xs), int(q))
    assert (ys.value, xs.value) == (hp, wp)
    return buf[:n].tobytes(), rcon


def run_job(job):
    key, work = job["key"], job["work"]
    out = os.path.join(work, key + ".json")
    if os.path.exists(out):
        return key, 0.0
    t = time.time()
    kind = job["kind"]
    if kind == "c3":
        img, lib, limit = W.config3_image(job["i"]), R.REF_SO, 8192
    elif kind == "c4":
        img, lib, limit = W.config4_image(job["i"]), R.REF_SO, 8192
    elif kind == "c5crop":
        img, lib, limit = W.config5_image(), R.REF_SO, 8192
    else:
        img, lib, limit = W.config5_image(), REF_XL, 16384
    s, r = _encode(lib, img, job["q"], limit)
    ent = {"in_shape": list(img.shape), "in_sha256": R.sha(img.tobytes()), "q": job["q"], "len": len(s),
           "stream_sha256": R.sha(s), "rcon_shape": list(r.shape), "rcon_sha256": R.sha(r.tobytes()),
           "cpu_seconds": round(time.time() - t, 1)}
    if kind != "c3":
        ent["rcon_band_sha"] = [R.sha(r[y:y + 32].tobytes())[:12] for y in range(0, r.shape[0], 32)]
    with open(out + ".tmp", "w") as f:
        json.dump(ent, f)
    os.replace(out + ".tmp", out)
    return key, time.time() - t


def merge(work):
    ents = {f[:-5]: json.load(open(os.path.join(work, f))) for f in sorted(os.listdir(work)) if f.endswith(".json")}
    c3 = {k.split("_")[1]: v for k, v in ents.items() if k.startswith("c3_")}
    c4 = {}
    for k, v in ents.items():
        if k.startswith("c4_"):
            _, i, q = k.split("_")
            c4.setdefault(i, {})[q] = v
    c5 = {k[3:]: v for k, v in ents.items() if k.startswith("c5_")}
    hdr = "the unmodified reference (oracle/_ref) on tests/workloads.py pictures; made by tools/make_golden_configs.py"
    for name, d in (("config3", c3), ("config4", c4), ("config5", c5)):
        if d:
            json.dump({"_source": hdr, "pictures": d}, open(os.path.join(GOLD, name + "_manifest.json"), "w"),
                      indent=1, sort_keys=True)
            print(name, len(d), "entries")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=6)
    ap.add_argument("--n3", type=int, default=256)
    ap.add_argument("--n4", type=int, default=8)
    ap.add_argument("--work", default="/tmp/hevce_golden_work")
    ap.add_argument("--merge", action="store_true")
    ap.add_argument("--no5", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.work, exist_ok=True)
    if a.merge:
        return merge(a.work)
    jobs = []
    if not a.no5:   # the longest jobs first
        jobs.append({"kind": "c5xl", "key": "c5_xl", "q": 2})
        jobs.append({"kind": "c5crop", "key": "c5_crop", "q": 2})
    for i in range(a.n4):
        for q in (0, 4):
            jobs.append({"kind": "c4", "key": f"c4_{i}_q{q}", "i": i, "q": q})
    for i in range(a.n3):
        jobs.append({"kind": "c3", "key": f"c3_{i:04d}", "i": i, "q": 2})
    for j in jobs:
        j["work"] = a.work
    with ProcessPoolExecutor(a.jobs) as ex:
        for key, dt in ex.map(run_job, jobs):
            print(f"{key} {dt:.0f}s", flush=True)
    merge(a.work)


if __name__ == "__main__":
    main()
