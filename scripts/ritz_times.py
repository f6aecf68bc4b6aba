"""Device times of the two Ritz-contraction kernels (FMA-pipe register tile vs DMMA tile) at the shapes of BASELINE cfg 4 / cfg 5.
   python scripts/ritz_times.py [n js d reps] ...   -> one JSON line per shape (gpurun_out/ritz_times.jsonl)"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdlibrs_b200 import _lib
L = _lib.lib()
L.svdlas2_test_time_ritz.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
L.svdlas2_test_time_ritz.restype = C.c_int
shapes = [(2_000_000, 480, 200, 2), (500_000, 2252, 1000, 1)]
if len(sys.argv) > 1:
    a = [int(x) for x in sys.argv[1:]]
    shapes = [tuple(a[i:i + 4]) for i in range(0, len(a), 4)]
out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "ritz_times.jsonl"), "a")
for n, js, d, reps in shapes:
    ms, diff, err = (C.c_double * 2)(), C.c_double(), C.c_double()
    rc = L.svdlas2_test_time_ritz(n, js, d, reps, -1, ms, C.byref(diff), C.byref(err))
    flops = 2.0 * n * js * d
    rec = dict(n=n, js=js, d=d, rc=rc, ms_fma=ms[0], ms_dmma=ms[1], tflops_fma=flops / ms[0] / 1e9 if ms[0] else None,
               tflops_dmma=flops / ms[1] / 1e9 if ms[1] else None, max_rel_diff=diff.value, max_rel_err_dmma=err.value)
    print("RITZ " + json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n"); out.flush()
