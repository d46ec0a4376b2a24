#!/usr/bin/env python
"""Input preparation for the CLI (the job of the reference's ConvertToPGM.py:35-85): any picture PIL can open ->
8-bit monochrome binary PGM (P5), the only input format `HEVCe` reads.

    python tools/convert_to_pgm.py <input_file>  <output_file(.pgm)>
    python tools/convert_to_pgm.py <input-dir>   <output-dir>

A single output name without the .pgm suffix gets it appended; in directory mode unreadable files are skipped.
"""
import os
import sys

import numpy as np


def to_pgm(src, dst):
    from PIL import Image
    with Image.open(src) as im:
        a = np.asarray(im.convert("L"))
    with open(dst, "wb") as f:                       # written by hand: exactly "P5\n<w> <h>\n255\n" + raster
        f.write(b"P5\n%d %d\n255\n" % (a.shape[1], a.shape[0]))
        f.write(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
    return a.shape


def main(argv):
    try:
        a, b = argv[1:3]
        assert a != b
    except Exception:
        print("\n    Usage :\n        python  %s  <input_file(.jpg|.png|.tiff|...)>  <output_file(.pgm)>\n    or :\n"
              "        python  %s  <input-dir>  <output-dir>\n" % (argv[0], argv[0]))
        return -1
    if not os.path.isdir(a):
        out = b if os.path.splitext(b)[1] == ".pgm" else b + ".pgm"
        try:
            to_pgm(a, out)
        except Exception:
            print("could not convert %s" % a)
            return -1
        return 0
    if not os.path.exists(b):
        print("mkdir %s\n" % b)
        os.mkdir(b)
    for fname in sorted(os.listdir(a)):
        out = os.path.join(b, os.path.splitext(fname)[0] + ".pgm")
        try:
            to_pgm(os.path.join(a, fname), out)
        except Exception:
            print("skip %s" % os.path.join(a, fname))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
