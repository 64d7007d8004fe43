import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymcl_b200
from pymcl_b200 import _lib
ctx = _lib.Context(0)
names = ["max3 (VIMNMX3)", "max2 (+1 LOP3)", "add3 (IADD3)", "mad (IMAD)", "fmaxf (+1 LOP3)", "max3 + mad pair", "vmaxs2 (+1 LOP3)"]
for op, n in enumerate(names):
    rate, ms = ctx.microbench(50 + op, 20000)
    print("%-20s %7.2f T ops/s = %5.1f lanes/clk/SM" % (n, rate / 1e12, rate / 148 / 1.965e9))
