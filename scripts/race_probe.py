import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVDLAS2_BLOCKED"] = "2"
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
A, _ = synthetic.make("cfg2", scale=float(sys.argv[1]) if len(sys.argv) > 1 else 0.02)
rng = np.random.default_rng(0)
with las2.Operator(A) as op:
    x = rng.standard_normal(A.shape[1])
    op.set("spmv_variant", 2); ref = op.opa(x)
    op.set("spmv_variant", 4)
    for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
        y = op.opa(x)
        bad = np.flatnonzero(np.abs(y - ref) > 1e-9 * (1 + np.abs(ref)))
        print("rep", rep, "bad", bad.size, bad[:6])
    op.set("spmv_variant", -1)
