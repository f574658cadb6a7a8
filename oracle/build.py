"""Compile the plain-C oracle (oracle/c/oracle.c) into oracle/_build/liboracle.so.
TEST INFRASTRUCTURE ONLY."""
import ctypes, os, subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "oracle.c")
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # no -ffast-math: the restatement is IEEE; the live reference (oracle/_ref) carries the
    # reference's own -ffast-math flag.
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", OUT, "-lm"])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        d = ctypes.c_double
        P = ctypes.c_void_p
        i64 = ctypes.c_int64
        ci = ctypes.c_int
        L.oracle_num_threads.restype = ci
        L.oracle_dft.argtypes = [P, P, i64, P, ci, ci, ci, d, d, d, P, P]
        L.oracle_dft.restype = None
        L.oracle_chisq.argtypes = [P, P, P, P, P, i64, ci]
        L.oracle_chisq.restype = ctypes.c_float
        L.oracle_loglike.argtypes = [P, P, P, P, P, i64, P, P, P]
        L.oracle_loglike.restype = d
        L.oracle_kernel.argtypes = [ci, d, d]
        L.oracle_kernel.restype = d
        L.oracle_grid_core.argtypes = [P, P, P, P, P, P, i64, ci, P, P, P, ci, d, P, P,
                                       ci, ci, d, ci, ci, ci, P, P, P]
        L.oracle_grid_core.restype = None
        _lib = L
    return _lib


if __name__ == "__main__":
    print(build(force=True))
