"""How does tcgen05.mma round its fp32 accumulator?  (GPU; pdsb_tc5_accum_probe)

One accumulator round of the tcgen05 DFT kernel (12 MMAs, K = 64, fp16 hi/lo operands) against the exact fp64 sum
of the same fp16 products, with the accumulator pre-loaded with 0 / +C / -C.  Prints, per operand pattern,
the mean and rms of the error in ulps (of the result for C = 0, of C otherwise)."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pdspy_b200 import _lib


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def run(a, b, c0, order):
    L = _lib.lib()
    ah, al = split(a)
    bh, bl = split(b)
    out = np.empty((128, 128), dtype=np.float32)
    p = lambda x: np.ascontiguousarray(x).view(np.uint16).ctypes.data_as(ctypes.c_void_p)
    ahc, alc, bhc, blc = (np.ascontiguousarray(x) for x in (ah, al, bh, bl))
    _lib.check(L.pdsb_tc5_accum_probe(p(ahc), p(alc), p(bhc), p(blc), ctypes.c_float(c0), order,
                                      out.ctypes.data_as(ctypes.c_void_p)))
    A_h, A_l, B_h, B_l = (x.astype(np.float64) for x in (ahc, alc, bhc, blc))
    exact = A_h @ B_h.T + A_h @ B_l.T + A_l @ B_h.T
    return out.astype(np.float64), exact


def patterns(rng, S):
    k = np.arange(64)
    yield "random sign", rng.uniform(-1, 1, (128, 64)), rng.uniform(-S, S, (128, 64))
    yield "all products > 0", rng.uniform(0, 1, (128, 64)), rng.uniform(0, S, (128, 64))
    yield "all products < 0", rng.uniform(0, 1, (128, 64)), rng.uniform(-S, 0, (128, 64))
    f = rng.uniform(0, 0.004, (128, 1))
    ph = rng.uniform(0, 1, (128, 1))
    prof = S * np.exp(-0.5 * ((k[None, :] - rng.uniform(0, 64, (128, 1))) / rng.uniform(5, 40, (128, 1))) ** 2)
    yield "low-frequency cos x smooth positive", np.cos(2 * np.pi * (f * k[None, :] + 0.1 * ph)), prof
    yield "cos x smooth positive, any frequency", np.cos(2 * np.pi * (rng.uniform(0, 0.5, (128, 1)) * k[None, :] + ph)), prof
    yield "small against C (x 2^-10)", rng.uniform(0, 1, (128, 64)), rng.uniform(0, S / 1024, (128, 64))


def main():
    rng = np.random.default_rng(7)
    S = 2.0 ** 15
    C = 1.5 * 2.0 ** 22
    ulpC = 2.0 ** (22 - 23)
    for order in (0, 1):
        print("== issue order %d (%s)" % (order, "product kernel" if order == 0 else "cross products first, hi.hi last"))
        for name, a, b in patterns(np.random.default_rng(7), S):
            out, exact = run(a, b, 0.0, order)
            ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(out), 1e-30))) - 23)
            e = (out - exact) / ulp * np.sign(exact)
            rel = (out - exact) / np.abs(exact)
            line = "%-40s C=0: err*sign/ulp(result) mean %+.3f rms %.3f  rel mean %+.2e |" % (name, e.mean(), e.std(), np.mean(rel * np.sign(exact)) if False else np.mean((out - exact) * np.sign(exact)) / np.mean(np.abs(exact)))
            for c0 in (C, -C):
                out, exact = run(a, b, c0, order)
                e = (out - c0 - exact) / ulpC
                line += "  C=%+.0e: mean %+.3f rms %.3f" % (c0, e.mean(), e.std())
            print(line)
    # does the offset error depend on where the data sits relative to ulp(C)?  scan the data scale
    print("== scale scan, order 0, C=+%g (error in ulp(C))" % C)
    for sh in (0, 2, 4, 6, 8, 12, 16):
        a = rng.uniform(0, 1, (128, 64))
        b = rng.uniform(0, S / 2.0 ** sh, (128, 64))
        out, exact = run(a, b, C, 0)
        e = (out - C - exact) / ulpC
        out2, exact2 = run(-a, b, C, 0)
        e2 = (out2 - C - exact2) / ulpC
        print("data scale 2^-%-2d  products>0: mean %+.3f rms %.3f   products<0: mean %+.3f rms %.3f" % (sh, e.mean(), e.std(), e2.mean(), e2.std()))


if __name__ == "__main__":
    main()
