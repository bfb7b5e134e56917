#!/usr/bin/env python
"""Role timelines of one CTA of the pipelined pair kernel (first ~6 items).  usage: pair_trace.py k [T B]"""
import ctypes as C, os, sys
here = os.path.dirname(os.path.abspath(__file__))
os.environ["NVSE_RB_TRACE"] = "1"; os.environ["NVSE_RB_T32"] = "1"
k = sys.argv[1]; T = sys.argv[2] if len(sys.argv) > 2 else "55168"; B = sys.argv[3] if len(sys.argv) > 3 else "32"
sys.argv = [os.path.join(here, "rb_bench.py"), "128", k, T, B, "1", "1"]
exec(open(os.path.join(here, "rb_bench.py")).read())
buf = (C.c_longlong * 128)()
lib.nvse_debug_rb_trace.argtypes = [C.POINTER(C.c_longlong)]
assert lib.nvse_debug_rb_trace(buf) == 0
rows = {"MMA   (issue start, issue end) per conv": buf[0:40], "worker(acc1 ready, epi1 end, acc2 ready, final end) per item": buf[40:80],
        "loader(start, buffer free, done) per item": buf[80:120]}
t0 = min(v for r in rows.values() for v in r if v)
for name, r in rows.items():
    print(name, [v - t0 for v in r if v][:28])
