#!/usr/bin/env python
"""Phase timeline of one CTA of the fused ResBlock kernel (run with NVSE_RB_TRACE=1).
usage: NVSE_RB_TRACE=1 rb_trace.py C k T B [npairs]"""
import ctypes as C
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
os.environ["NVSE_RB_TRACE"] = "1"
sys.argv = [os.path.join(here, "rb_bench.py")] + sys.argv[1:] + (["3"] if len(sys.argv) < 6 else []) + ["1"]
exec(open(os.path.join(here, "rb_bench.py")).read())
buf = (C.c_longlong * 128)()
lib.nvse_debug_rb_trace.argtypes = [C.POINTER(C.c_longlong)]
assert lib.nvse_debug_rb_trace(buf) == 0
print('cycles the MMA warp spent waiting for weight stages:', buf[63])
mma = [v for v in buf[:63] if v]
wrk = [v for v in buf[64:] if v]
t0 = min(mma + wrk)
print("MMA warp  (wait-start, wait-end per conv, ..., done):", [v - t0 for v in mma])
print("worker 0  (load start, load end, [acc ready, epi1 end, x ready, epi2 end]*, done):", [v - t0 for v in wrk])
