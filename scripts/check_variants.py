"""Cross-check SpMV variants against each other at full size (no oracle: variant 2/3 are parity-tested).
   python scripts/check_variants.py [--workload cfg2] [--scale 1.0] [--blocked 2]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--blocked", default="2")
a = ap.parse_args()
os.environ["SVDLAS2_BLOCKED"] = a.blocked
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
A, _ = synthetic.make(a.workload, scale=a.scale)
rng = np.random.default_rng(0)
with las2.Operator(A) as op:
    for transposed in (False, True):
        x = rng.standard_normal(A.shape[0] if transposed else A.shape[1])
        ys = {}
        for v, hot in [(2, -1), (3, -1), (4, -1), (3, -1), (4, -1), (3, 0), (4, 0), (3, -1), (4, 4096)]:
            op.set("spmv_variant", v); op.set("spmv_hot", hot)
            ys[(v, hot, len(ys))] = op.opa(x, transposed)
        ref = ys[(2, -1, 0)]
        for k, y in ys.items():
            err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
            bad = np.flatnonzero(np.abs(y - ref) > 1e-9 * (1 + np.abs(ref)))
            print(f"transposed={transposed} variant={k}: rel err {err:.3e}, bad entries {bad.size}" + (f" first {bad[:8]} y={y[bad[:4]]} ref={ref[bad[:4]]}" if bad.size else ""))
    op.set("spmv_variant", -1); op.set("spmv_hot", -1)
