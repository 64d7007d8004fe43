"""int8 tensor-pipe peak (mcl_microbench 0) as a function of how long the kernel runs: burst vs sustained (power cap)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymcl_b200
from pymcl_b200 import _lib
ctx = _lib.default_context()
for it in (4000, 20000, 100000, 400000, 400000, 4000):
    rate, ms = ctx.microbench(0, it)
    print(json.dumps({"iters": it, "ms": round(ms, 3), "POPs": round(rate / 1e15, 4)}))
