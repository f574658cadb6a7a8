"""interpolate_model(code="nufft") on the C3 cube: time per call with the result left on the device (GPU).
PDSB_NUFFT_DIRECT=1 times the sampler without shared-memory patches for comparison."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pdspy_b200 as pb
import synth
from pdspy_b200 import _lib
from pdspy_b200.interferometry import interpolate_model
pb.set_cache_policy("freeze")
c = synth.make_config("C3")
for code in ("nufft", "galario-fft" if len(sys.argv) > 1 else "nufft"):
    for _ in range(2):
        m = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"], code=code)
    _lib.check(_lib.lib().pdsb_synchronize())
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        m = interpolate_model(c["u"], c["v"], c["freq"], c["model"], dRA=c["dRA"], dDec=c["dDec"], code=code)
    _lib.check(_lib.lib().pdsb_synchronize())
    print("code=%s: %.2f ms per call (pageable 134 MB cube in, [1M, 64] visibilities left on the device)" % (code, (time.perf_counter() - t0) / n * 1e3))
    print("  checksum", float(m.real[::1000].sum()))
