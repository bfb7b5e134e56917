#!/usr/bin/env python
"""Timeline of one CTA of the software-pipelined chain kernel.  usage: pipe_trace.py C k [T B]"""
import ctypes as C, os, sys
here = os.path.dirname(os.path.abspath(__file__))
os.environ["NVSE_RB_TRACE"] = "1"; os.environ["NVSE_RB_T32"] = "1"
Cc, k = sys.argv[1], sys.argv[2]
T = sys.argv[3] if len(sys.argv) > 3 else str(110336 if Cc == "64" else 220672); B = sys.argv[4] if len(sys.argv) > 4 else "32"
if Cc == "32": os.environ["NVSE_RB_H16"] = "1"
sys.argv = [os.path.join(here, "rb_bench.py"), Cc, k, T, B, "3", "1"]
exec(open(os.path.join(here, "rb_bench.py")).read())
buf = (C.c_longlong * 128)()
lib.nvse_debug_rb_trace.argtypes = [C.POINTER(C.c_longlong)]
assert lib.nvse_debug_rb_trace(buf) == 0
mma = [v for v in buf[0:60] if v]; wrk = [v for v in buf[64:124] if v]
t0 = min(mma + wrk)
print("MMA  (issue start, issue end) per (conv, half):", [v - t0 for v in mma])
print("work (load start, load end, then acc-ready per (conv, half), ..., done):", [v - t0 for v in wrk])
