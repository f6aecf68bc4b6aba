"""Device-time the purge reorthogonalisation pass (c = Q'[q r], then [q r] -= Q c) on synthetic panels of the BASELINE
shapes and report achieved HBM GB/s against 2 * ncols * N * 8 bytes per pass.
   python scripts/panel_bench.py [--shapes 100000x130,500000x1150,2000000x300]"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdlibrs_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="100000x130,200000x130,500000x1150,500000x300,2000000x300")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
L = _lib.lib()
L.svdlas2_test_time_panel.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
L.svdlas2_test_time_panel.restype = C.c_int
for sh in a.shapes.split(","):
    n, k = (int(v) for v in sh.split("x"))
    d, u = C.c_double(), C.c_double()
    rc = L.svdlas2_test_time_panel(n, k, a.reps, 0, C.byref(d), C.byref(u))
    assert rc == 0, rc
    half = n * k * 8 / 1e6  # MB read per kernel
    print(json.dumps(dict(n=n, ncols=k, ms_dots=round(d.value, 4), ms_update=round(u.value, 4), GBs_dots=round(half / d.value, 1),
                          GBs_update=round(half / u.value, 1), GBs_pass=round(2 * half / (d.value + u.value), 1))), flush=True)
