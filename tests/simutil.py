"""ctypes wrapper for the host-compiled kernel simulator (tests/sim, TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "sim")
SIM_SO = os.path.join(SIM_DIR, "libhevce_sim.so")
CSRC = os.path.join(ROOT, "hevc-image-encoder-lite_b200", "csrc")
_u8p = ctypes.POINTER(ctypes.c_ubyte)


def build_sim(force=False):
    srcs = [os.path.join(SIM_DIR, "hevce_sim.cpp"), os.path.join(CSRC, "hevce_core.h"), os.path.join(CSRC, "hevce_xform_gen.h")]
    if not force and os.path.exists(SIM_SO) and all(os.path.getmtime(SIM_SO) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", SIM_SO, srcs[0]], check=True)


_lib = None


def sim():
    global _lib
    if _lib is None:
        build_sim()
        _lib = ctypes.CDLL(SIM_SO)
        _lib.hevce_sim_encode.restype = ctypes.c_int
    return _lib


def sim_encode(img, q, order=0, max_dim=8192):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    hp, wp = (min(h, max_dim) + 31) // 32 * 32, (min(w, max_dim) + 31) // 32 * 32
    rcon = np.zeros((hp, wp), np.uint8)
    cap = 256 + 2 * hp * wp
    out = np.zeros(cap, np.uint8)
    ys, xs, err = ctypes.c_int(h), ctypes.c_int(w), ctypes.c_int(0)
    n = sim().hevce_sim_encode(out.ctypes.data_as(_u8p), cap, img.ctypes.data_as(_u8p), rcon.ctypes.data_as(_u8p),
                               ctypes.byref(ys), ctypes.byref(xs), int(q), int(order), int(max_dim), ctypes.byref(err))
    assert (ys.value, xs.value) == (hp, wp)
    return out[:n].tobytes(), rcon, err.value


def sim_last_partition(shape):
    """(cu_size, mode, kind) maps of the last sim_encode call; shape = padded picture shape."""
    hp, wp = shape
    cu, mode, kind = np.zeros((hp // 4, wp // 4), np.uint8), np.zeros((hp // 4, wp // 4), np.uint8), np.zeros((hp // 8, wp // 8), np.uint8)
    sim().hevce_sim_last_partition(cu.ctypes.data_as(_u8p), mode.ctypes.data_as(_u8p), kind.ctypes.data_as(_u8p))
    return cu, mode, kind
