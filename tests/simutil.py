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


# ---- gang simulator: one host thread per picture of a gang (tests/sim/hevce_simgang.cpp)
SIMGANG_SO = os.path.join(SIM_DIR, "libhevce_simgang.so")
_glib = None


def build_simgang(force=False):
    srcs = [os.path.join(SIM_DIR, "hevce_simgang.cpp"), os.path.join(CSRC, "hevce_core.h"), os.path.join(CSRC, "hevce_xform_gen.h")]
    if not force and os.path.exists(SIMGANG_SO) and all(os.path.getmtime(SIMGANG_SO) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", CSRC, "-o", SIMGANG_SO, srcs[0]], check=True)


def simgang():
    global _glib
    if _glib is None:
        build_simgang()
        _glib = ctypes.CDLL(SIMGANG_SO)
        _glib.hevce_simgang_encode.restype = ctypes.c_int
        _glib.hevce_simgang_size.restype = ctypes.c_int
    return _glib


def simgang_encode(imgs, qs, order=0):
    """Encode up to GANG pictures of identical size as one gang. Returns [(stream, rcon, err), ...]."""
    imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in imgs]
    n = len(imgs)
    h, w = imgs[0].shape
    assert all(i.shape == (h, w) for i in imgs)
    hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    cap = 256 + 2 * hp * wp
    outs = [np.zeros(cap, np.uint8) for _ in range(n)]
    rcons = [np.zeros((hp, wp), np.uint8) for _ in range(n)]
    arr = lambda xs: (_u8p * n)(*[x.ctypes.data_as(_u8p) for x in xs])
    qa = (ctypes.c_int * n)(*[int(q) for q in qs])
    lens, errs = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    rc = simgang().hevce_simgang_encode(n, arr(outs), cap, arr(imgs), arr(rcons), h, w, qa, int(order), lens, errs)
    assert rc == 0
    return [(outs[i][: lens[i]].tobytes(), rcons[i], errs[i]) for i in range(n)]
