"""Warp-chunk SpMV at full size: agreement with the parity-tested tile kernel, then a device-timed sweep (CUDA events)
over the shared-memory prefix size.   python scripts/chunk_check.py [--workload cfg2] [--scale 1.0] [--hots 0,2048,4096,8192]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--hots", default="-1,0,2048,4096,8192"); ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--pads", default="0"); ap.add_argument("--carves", default="-1"); ap.add_argument("--shapes", default="32"); ap.add_argument("--pdls", default="1,0")
ap.add_argument("--old", action="store_true", help="also time the previous default (tile stream) for comparison")
a = ap.parse_args()
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
A, dims = synthetic.make(a.workload, scale=a.scale)
print(json.dumps(dict(workload=a.workload, shape=list(A.shape), nnz=int(A.nnz))), flush=True)
rng = np.random.default_rng(0)
with las2.Operator(A) as op:
    for transposed in (False, True):
        x = rng.standard_normal(A.shape[0] if transposed else A.shape[1])
        op.set("spmv_variant", 2); ref = op.opa(x, transposed)
        for hot, shape in ((-1, 32), (0, 32), (100, 16), (-1, 16)):
            op.set("spmv_variant", 5); op.set("spmv_hot", hot); op.set("spmv_shape", shape)
            for rep in range(2):
                y = op.opa(x, transposed)
                bad = np.flatnonzero(~(np.abs(y - ref) <= 1e-11 * np.abs(ref).max()))
                print(f"transposed={transposed} hot={hot} shape={shape} rep={rep}: max err {np.abs(y - ref).max():.3e} (scale {np.abs(ref).max():.3e}), bad {bad.size}"
                      + (f" first {bad[:8]} y={y[bad[:4]]} ref={ref[bad[:4]]}" if bad.size else ""), flush=True)
    for hot, pad, carve, shape, pdl in [(int(h), int(p), int(c), int(sh), int(pd)) for h in a.hots.split(",") for p in a.pads.split(",") for c in a.carves.split(",") for sh in a.shapes.split(",") for pd in a.pdls.split(",")]:
        op.set("spmv_variant", 5); op.set("spmv_hot", hot); op.set("spmv_smem_pad_kb", pad); op.set("spmv_carveout", carve); op.set("spmv_shape", shape); op.set("spmv_pdl", pdl)
        t = op.time_opb(reps=a.reps, flush_l2=False)
        print(json.dumps(dict(variant="chunk", hot=hot, pad_kb=pad, carve=carve, shape=shape, pdl=pdl, ms_opb=round(t["ms_per_opb"], 4), ms_b=round(t["ms_spmv_b"], 4),
                              ms_bt=round(t["ms_spmv_bt"], 4), GBs_opb=round(t["bytes_opb"] / t["ms_per_opb"] / 1e6, 1),
                              GBs_b=round(t["bytes_spmv_b"] / t["ms_spmv_b"] / 1e6, 1),
                              GBs_bt=round(t["bytes_spmv_bt"] / t["ms_spmv_bt"] / 1e6, 1))), flush=True)
    if a.old:
        for v in (3, 2):
            op.set("spmv_variant", v); op.set("spmv_hot", -1)
            t = op.time_opb(reps=a.reps, flush_l2=False)
            print(json.dumps(dict(variant=v, ms_opb=round(t["ms_per_opb"], 4), ms_b=round(t["ms_spmv_b"], 4), ms_bt=round(t["ms_spmv_bt"], 4))), flush=True)
    op.set("spmv_variant", -1); op.set("spmv_hot", -1); op.set("spmv_smem_pad_kb", 0); op.set("spmv_carveout", -1); op.set("spmv_shape", 32); op.set("spmv_pdl", -1)
    import time
    for i in range(2):
        t0 = time.perf_counter(); r = op.solve(dims, random_seed=12345); dt = time.perf_counter() - t0
        print("solve", i, round(dt, 4), {k: round(v, 4) for k, v in r.profile.items() if k.startswith("t_")}, "steps", r.diagnostics.lanczos_steps, flush=True)
