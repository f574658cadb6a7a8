import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdspy_b200 import _lib
L = _lib.lib()
names = {0: "FFMA reg chain", 1: "FFMA2 reg chain", 2: "pattern uv2 regs", 3: "pattern uv2 LDS.128 bcast",
         4: "pattern uv4 LDS", 5: "pattern uv1 LDS", 6: "pattern uv4 regs", 7: "pattern uv3 LDS",
         8: "const-bank uv2 tp8", 9: "const-bank uv4 tp8", 10: "const-bank uv2 tp16", 11: "mma.sync m16n8k8 tf32 (legacy TC)", 12: "mma.sync m16n8k16 f16 (legacy TC)", 13: "DFMA reg chain (fp64)"}
for var in (1, 13):
    tf, ms = ctypes.c_double(), ctypes.c_double()
    for rep in range(2):
        _lib.check(L.pdsb_bench_fma(var, 20000 if var != 13 else 500, ctypes.byref(tf), ctypes.byref(ms)))
    print("fma bench %d %-28s %.2f TFLOP/s (%.1f%% of 74.45) %.2f ms" % (var, names[var], tf.value, tf.value / 74.45 * 100, ms.value), flush=True)
