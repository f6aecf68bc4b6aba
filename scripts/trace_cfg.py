import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVDLAS2_TRACE"] = "1"
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
name, scale = sys.argv[1], float(sys.argv[2])
A, dims = synthetic.make(name, scale=scale)
with las2.Operator(A) as op:
    keep = []
    for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
        t0 = time.perf_counter(); r = op.solve(dims, random_seed=12345); print("solve", i, round(time.perf_counter() - t0, 3), {k: round(v, 3) for k, v in r.profile.items() if k.startswith("t_")}, flush=True)
        (keep.append(r) if len(sys.argv) > 4 else None); r = None
