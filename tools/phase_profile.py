#!/usr/bin/env python
"""Per-phase latency histogram (development build `make -C hevc-image-encoder-lite_b200/csrc profile`).
usage: python tools/phase_profile.py [n] [h] [w] [qpd6]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hevc-image-encoder-lite_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hevce_b200 as H
H.LIB_PATH = os.path.join(ROOT, "hevc-image-encoder-lite_b200", "libhevce_b200_prof.so")
import workloads as WL
n, h, w, q = (int(v) for v in (sys.argv[1:] + ["1036", "64", "64", "2"][len(sys.argv) - 1:]))
imgs = [WL.config3_image(i)[100:100 + h, 200:200 + w].copy() for i in range(n)]
ses = H.Session(0, [i.shape for i in imgs], q)
ses.upload(imgs)
print("ms", ses.encode())
H.lib().hevce_profile_dump()
