"""DFMA latency / ILP probe on the box: TFLOP/s for (chains per thread, warps per SM)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

g.build()
from jstacs_b200 import _native

L = _native.lib()
v = C.c_double()
print("chains threads blocks/SM warps/SM TFLOP/s")
for threads, bps in ((128, 1), (256, 1), (384, 1), (512, 1), (512, 2), (512, 4)):
    for chains in (1, 2, 4, 8, 16):
        _native.check(L.jh_probe_fp64(chains, threads, bps, C.byref(v)))
        print(f"{chains:6d} {threads:7d} {bps:9d} {threads * bps // 32:8d} {v.value:8.2f}")

print("DMMA m8n8k4: chains threads blocks/SM warps/SM TFLOP/s")
for threads, bps in ((128, 1), (256, 1), (512, 1), (512, 2)):
    for chains in (1, 2, 4, 8):
        _native.check(L.jh_probe_dmma(chains, threads, bps, C.byref(v)))
        print(f"{chains:6d} {threads:7d} {bps:9d} {threads * bps // 32:8d} {v.value:8.2f}")
