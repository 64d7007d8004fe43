import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymcl_b200
from pymcl_b200 import _lib
ctx = _lib.Context(0)
iters = 4000
print("variant -> TOP/s, cycles per MMA @1.965GHz")
for n in (64, 128, 256):
    for ts in (0, 1):
        for ext in (0, 1):
            which = 100 + n // 64 + 10 * ts + 100 * ext
            rate, ms = ctx.microbench(which, iters)
            cyc = ms * 1e-3 * 1.965e9 / (iters * 15)
            print("n=%3d ts=%d ext=%d : %7.1f TOP/s  %6.1f cyc/MMA (ideal %d)" % (n, ts, ext, rate / 1e12, cyc, n // 2))
