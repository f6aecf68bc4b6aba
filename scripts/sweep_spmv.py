"""SpMV variant sweep on one workload (device-timed opb, CUDA events).
   python scripts/sweep_spmv.py [--workload cfg2] [--scale 1.0]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--blocked", default="0")
a = ap.parse_args()
os.environ['SVDLAS2_BLOCKED'] = a.blocked
A, dims = synthetic.make(a.workload, scale=a.scale)
print(json.dumps(dict(workload=a.workload, shape=list(A.shape), nnz=int(A.nnz))))
with las2.Operator(A) as op:
    for variant, hot in ([(4, -1), (4, 1024), (4, 2048), (4, 3072), (4, 6144), (3, -1)] if a.blocked != '0' else [(3, -1), (2, -1)]):
        op.set("spmv_variant", variant); op.set("spmv_hot", hot)
        t = op.time_opb(reps=a.reps, flush_l2=False)
        print(json.dumps(dict(variant=variant, hot=hot, ms_opb=round(t["ms_per_opb"], 4), ms_b=round(t["ms_spmv_b"], 4),
                              ms_bt=round(t["ms_spmv_bt"], 4), GBs_opb=round(t["bytes_opb"] / t["ms_per_opb"] / 1e6, 1),
                              GBs_b=round(t["bytes_spmv_b"] / t["ms_spmv_b"] / 1e6, 1),
                              GBs_bt=round(t["bytes_spmv_bt"] / t["ms_spmv_bt"] / 1e6, 1))), flush=True)
    op.set("spmv_variant", -1); op.set("spmv_hot", -1)
