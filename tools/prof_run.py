#!/usr/bin/env python
"""Small fixed workload for ncu: N pictures of HxW cut from the config-3 generator, encoded `reps` times in one
device-resident session.  usage: python tools/prof_run.py [n] [h] [w] [qpd6] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hevc-image-encoder-lite_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hevce_b200 as H  # noqa: E402
import workloads as WL  # noqa: E402

n, h, w, q, reps = (int(v) for v in (sys.argv[1:] + ["148", "64", "64", "2", "2"][len(sys.argv) - 1:]))
oy, ox = min(100, 512 - h), min(200, 768 - w)   # full-size pictures: the bench workload itself
imgs = [WL.config3_image(i)[oy:oy + h, ox:ox + w].copy() for i in range(n)]
ses = H.Session(0, [i.shape for i in imgs], q)
ses.upload(imgs)
for r in range(reps):
    ms = ses.encode()
    print(f"rep {r}: {ms:.2f} ms, {n * h * w / ms / 1e3:.3f} Mpx/s, grid {ses.grid}")
streams, _ = ses.download()
print("bytes", sum(len(s) for s in streams))
