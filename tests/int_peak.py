"""Integer-pipe micro-benchmark: lane-ops/clk/SM per SASS instruction class, on cuda:0."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqalib_b200 import capi

NAMES = ["IADD3", "VIMNMX.S32", "VIADDMNMX.S32", "VIMNMX3.S16x2", "IMAD", "LOP3", "PRMT", "VIADD.16x2",
         "VIADDMNMX.S16x2.RELU", "packed SW cell mix (5 ops)"]
lib = capi.Lib()
out = {}
for k, name in enumerate(NAMES):
    v, mhz = C.c_double(), C.c_double()
    lib.check(lib.L.seqa_cuda_int_peak(0, k, C.byref(v), C.byref(mhz)))
    out[name] = {"lane_ops_per_clk_per_sm": round(v.value, 2), "sm_mhz": round(mhz.value, 1)}
    print("%-28s %7.2f lane-ops/clk/SM  @ %.0f MHz" % (name, v.value, mhz.value))
print(json.dumps(out))
