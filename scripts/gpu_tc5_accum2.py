"""Lattice split of the tcgen05 operands (GPU; pdsb_tc5_accum_probe): hi limbs on a fixed-point lattice so coarse
that the running sum of the hi.hi products is exactly representable in the fp32 accumulator; cross products
first.  Expected: one effective truncation per accumulator round (-0.5 ulp of the result) instead of ~12."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pdspy_b200 import _lib


def probe(ah, al, bh, bl, c0, order):
    L = _lib.lib()
    out = np.empty((128, 128), dtype=np.float32)
    arrs = [np.ascontiguousarray(x.astype(np.float16)) for x in (ah, al, bh, bl)]
    p = lambda x: x.view(np.uint16).ctypes.data_as(ctypes.c_void_p)
    _lib.check(L.pdsb_tc5_accum_probe(p(arrs[0]), p(arrs[1]), p(arrs[2]), p(arrs[3]), ctypes.c_float(c0), order,
                                      out.ctypes.data_as(ctypes.c_void_p)))
    A_h, A_l, B_h, B_l = (x.astype(np.float64) for x in arrs)
    exact3 = A_h @ B_h.T + A_h @ B_l.T + A_l @ B_h.T
    return out.astype(np.float64), exact3


def split_native(x):
    hi = x.astype(np.float16).astype(np.float64)
    return hi, (x - hi).astype(np.float16).astype(np.float64)


def split_lattice(x, q):
    hi = np.rint(x / q) * q
    return hi, (x - hi).astype(np.float16).astype(np.float64)


def report(name, out, exact3, true):
    ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(out), 1e-30))) - 23)
    e = (out - exact3) / ulp * np.sign(exact3)
    big = np.abs(true) > 0.25 * np.abs(true).max()
    print("%-58s accum err*sign/ulp mean %+.3f rms %.3f | rel bias (accum) %+.2e  (vs true a.b) %+.2e  rms rel err vs true (large elems) %.2e"
          % (name, e.mean(), e.std(), np.mean((out - exact3) * np.sign(exact3)) / np.mean(np.abs(exact3)),
             np.mean((out - true) * np.sign(true)) / np.mean(np.abs(true)),
             np.sqrt(np.mean(((out - true)[big] / true[big]) ** 2))))


def main():
    rng = np.random.default_rng(11)
    k = np.arange(64)
    cases = []
    cases.append(("all products > 0, flat", rng.uniform(0, 1, (128, 64)), rng.uniform(0, 1, (128, 64))))
    f = rng.uniform(0, 0.004, (128, 1))
    prof = np.exp(-0.5 * ((k[None, :] - rng.uniform(0, 64, (128, 1))) / rng.uniform(5, 40, (128, 1))) ** 2)
    cases.append(("low-frequency cos x smooth positive", np.cos(2 * np.pi * (f * k[None, :] + 0.02)), prof))
    star = 1e-5 * prof.copy()
    star[:, 31] = 1.0
    cases.append(("star pixel 1e5 x disk", np.cos(2 * np.pi * (f * k[None, :] + 0.02)), star))
    cases.append(("random sign", rng.uniform(-1, 1, (128, 64)), rng.uniform(-1, 1, (128, 64))))
    cases.append(("cos any frequency x smooth positive", np.cos(2 * np.pi * (rng.uniform(0, 0.5, (128, 1)) * k[None, :] + rng.uniform(0, 1, (128, 1)))), prof))
    for name, a, b in cases:
        true = a @ b.T
        # today's split: native fp16 rounding, image scaled so that max -> 2^14
        s = 2.0 ** np.floor(np.log2(16384.0 / np.abs(b).max()))
        ah, al = split_native(a)
        bh, bl = split_native(b * s)
        for order in (0, 1):
            out, ex = probe(ah, al, bh, bl, 0.0, order)
            report("%s | native split, order %d" % (name, order), out / s, ex / s, true)
        # lattice split: a on 2^-8, b on integers with max <= 2047 and max row L1 <= 2^16 - 64
        for abits in (8, 9):
            l1 = np.abs(b).sum(axis=1).max()
            lim = 2.0 ** (24 - abits)
            q = 2.0 ** np.ceil(np.log2(max(np.abs(b).max() / 2047.0, l1 / (lim - 64.0))))
            ah, al = split_lattice(a, 2.0 ** -abits)
            bh, bl = split_lattice(b / q, 1.0)
            for order in (1, 0):
                out, ex = probe(ah, al, bh, bl, 0.0, order)
                report("%s | lattice a 2^-%d, order %d" % (name, abits, order), out * q, ex * q, true)
        print()


if __name__ == "__main__":
    main()
