#!/usr/bin/env python
"""Generate hevc-image-encoder-lite_b200/csrc/hevce_xform_gen.h: straight-line 1-D HEVC core transforms.

The reference multiplies by the full NxN integer matrices (HEVCe.c:469-516, tables :391-464).  All sums fit int32,
so any exact regrouping gives identical results.  We emit even/odd partial butterflies (N -> N/2 recursion, the
odd half is an N/2 x N/2 matrix product) with every coefficient as a literal, so the compiler issues IMAD with
immediates and keeps all operands in registers.  MAC counts: 4:16(DST, dense) 8:22 16:86 32:342.

The matrices are generated from the 31 HEVC core-transform constants through the cosine symmetries and verified
here with numpy against a dense product on random input before the header is written.
"""
import os

import numpy as np

Q = [64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
     61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0]


def cosA(j):
    j &= 127
    if j > 64:
        j = 128 - j
    return -Q[64 - j] if j > 32 else Q[j]


def dct(n):
    step = 32 // n
    return np.array([[cosA(k * step * (2 * i + 1)) for i in range(n)] for k in range(n)], dtype=np.int64)


DST4 = np.array([[29, 55, 74, 84], [74, 74, 0, -74], [84, -29, -74, 55], [55, -84, 74, -29]], dtype=np.int64)


class Emit:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def let(self, expr):
        t = self.tmp()
        self.lines.append(f"    const int {t} = {expr};")
        return t


def lin(coefs, vars_):
    """sum coef*var as a C expression, skipping zeros"""
    terms = []
    for c, v in zip(coefs, vars_):
        c = int(c)
        if c == 0:
            continue
        terms.append(f"{c} * {v}")
    return " + ".join(terms).replace("+ -", "- ") if terms else "0"


def gen_fwd(e, n, xs):
    """returns list of n expressions/temps: y[k] = sum_i C_n[k][i] x[i]"""
    C = dct(n)
    if n == 4:
        e0 = e.let(f"{xs[0]} + {xs[3]}")
        e1 = e.let(f"{xs[1]} + {xs[2]}")
        o0 = e.let(f"{xs[0]} - {xs[3]}")
        o1 = e.let(f"{xs[1]} - {xs[2]}")
        return [e.let(f"64 * ({e0} + {e1})"), e.let(lin([C[1][0], C[1][1]], [o0, o1])),
                e.let(f"64 * ({e0} - {e1})"), e.let(lin([C[3][0], C[3][1]], [o0, o1]))]
    h = n // 2
    ev = [e.let(f"{xs[i]} + {xs[n - 1 - i]}") for i in range(h)]
    od = [e.let(f"{xs[i]} - {xs[n - 1 - i]}") for i in range(h)]
    even = gen_fwd(e, h, ev)
    out = [None] * n
    for k in range(h):
        out[2 * k] = even[k]
        out[2 * k + 1] = e.let(lin(C[2 * k + 1][:h], od))
    return out


def gen_inv(e, n, ys):
    """returns x[i] = sum_k C_n[k][i] y[k]"""
    C = dct(n)
    if n == 4:
        e0 = e.let(f"64 * ({ys[0]} + {ys[2]})")
        e1 = e.let(f"64 * ({ys[0]} - {ys[2]})")
        o0 = e.let(lin([C[1][0], C[3][0]], [ys[1], ys[3]]))
        o1 = e.let(lin([C[1][1], C[3][1]], [ys[1], ys[3]]))
        return [e.let(f"{e0} + {o0}"), e.let(f"{e1} + {o1}"), e.let(f"{e1} - {o1}"), e.let(f"{e0} - {o0}")]
    h = n // 2
    even = gen_inv(e, h, [ys[2 * k] for k in range(h)])
    out = [None] * n
    for i in range(h):
        o = e.let(lin([C[2 * k + 1][i] for k in range(h)], [ys[2 * k + 1] for k in range(h)]))
        out[i] = e.let(f"{even[i]} + {o}")
        out[n - 1 - i] = e.let(f"{even[i]} - {o}")
    return out


def func(name, n, body_fn):
    e = Emit()
    xs = [f"x[{i}]" for i in range(n)]
    outs = body_fn(e, n, xs)
    src = [f"HEVCE_HD void {name}(const int (&x)[{n}], int (&y)[{n}]) {{"] + e.lines
    src += [f"    y[{k}] = {o};" for k, o in enumerate(outs)] + ["}", ""]
    return "\n".join(src)


def dst_func(name, inverse):
    M = DST4.T if inverse else DST4
    lines = [f"HEVCE_HD void {name}(const int (&x)[4], int (&y)[4]) {{"]
    for k in range(4):
        lines.append(f"    y[{k}] = {lin(M[k], [f'x[{i}]' for i in range(4)])};")
    lines += ["}", ""]
    return "\n".join(lines)


def py_eval(src, name, n, x):
    """execute the generated C (it is also valid Python after trivial rewriting) to verify it"""
    body = src.split("{", 1)[1].rsplit("}", 1)[0]
    env = {"x": [int(v) for v in x], "y": [0] * n}
    for ln in body.strip().splitlines():
        ln = ln.strip().rstrip(";").replace("const int ", "")
        exec(ln, {}, env)
    return np.array(env["y"], dtype=np.int64)


def main():
    rng = np.random.default_rng(0)
    parts = []
    for n in (8, 16, 32):
        for inverse in (False, True):
            name = f"{'idct' if inverse else 'fdct'}{n}"
            src = func(name, n, gen_inv if inverse else gen_fwd)
            for _ in range(20):
                x = rng.integers(-32768, 32768, n)
                want = (dct(n).T if inverse else dct(n)) @ x
                got = py_eval(src, name, n, x)
                assert np.array_equal(want, got), name
            parts.append(src)
    for inverse in (False, True):
        name = "idst4" if inverse else "fdst4"
        src = dst_func(name, inverse)
        for _ in range(20):
            x = rng.integers(-32768, 32768, 4)
            assert np.array_equal((DST4.T if inverse else DST4) @ x, py_eval(src, name, 4, x))
        parts.append(src)
    hdr = ("// GENERATED by tools/gen_transforms.py -- do not edit.\n"
           "// 1-D HEVC core transforms as partial butterflies with literal coefficients (exact in int32).\n"
           "// fdctN: y = C_N x ; idctN: y = C_N^T x ; fdst4 / idst4: the 4x4 DST-VII pair.\n"
           "// Replaces the dense matMul of the reference (HEVCe.c:469-492, matrices :391-464).\n"
           "#pragma once\n\n")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "hevc-image-encoder-lite_b200", "csrc", "hevce_xform_gen.h")
    open(out, "w").write(hdr + "\n".join(parts))
    print("wrote", out)


if __name__ == "__main__":
    main()
