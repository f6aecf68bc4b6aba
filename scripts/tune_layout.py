"""Layout tuning on one GPU: generate a BASELINE workload ONCE, upload it under different layout knobs and device-time opb.
   python scripts/tune_layout.py [--workload cfg4] [--shard G]   -> gpurun_out/tune_layout.jsonl
Knobs swept: SVDLAS2_SPLIT_PER_PIECE (heavy/light cut of B'), SVDLAS2_SPLIT, SVDLAS2_WIDE_SLABS, SVDLAS2_RELABEL, spmv_hot."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--shard", type=int, default=1, help="time the first of this many nnz-balanced row shards (multi-GPU shard shape)")
ap.add_argument("--sets", default="base,pp3,pp4,pp8,pp12,nosplit,nowide,norelabel,hot12k,hot16k")
a = ap.parse_args()
from svdlibrs_b200 import synthetic, sharding
nrows, ncols, nnz_req, dims = synthetic.config_dims(a.workload, a.scale)
try:
    cores = len(os.sched_getaffinity(0))
except Exception:
    cores = os.cpu_count() or 1
rows = None
if a.shard > 1:
    b = sharding.partition_rows_by_nnz(synthetic.requested_row_offsets(a.workload, a.scale), a.shard)
    rows = (int(b[0]), int(b[1]))
t0 = time.time()
indptr, idx, val = synthetic.powerlaw_arrays(nrows, ncols, nnz_req, 12345, rows=rows, workers=min(32, cores))
print(f"generated {len(val)} nnz in {time.time() - t0:.1f} s", flush=True)
import svdlibrs_b200 as las2
m = (rows[1] - rows[0]) if rows else nrows
A = las2.RawMatrix("csr", (m, ncols), indptr.astype(np.uint64), idx.view(np.uint64), val)


def make_op():
    # a shard is uploaded exactly as a rank of a sharded run uploads it (B = the shard's rows, N side = all columns)
    if a.shard > 1:
        return las2.ShardedOperator(A, nrows, rows[0], int(len(val)) * a.shard, None, False)
    return las2.Operator(A)
SETS = {
    "base": ({}, {}),
    "pp3": ({"SVDLAS2_SPLIT_PER_PIECE": "3"}, {}), "pp4": ({"SVDLAS2_SPLIT_PER_PIECE": "4"}, {}),
    "pp8": ({"SVDLAS2_SPLIT_PER_PIECE": "8"}, {}), "pp12": ({"SVDLAS2_SPLIT_PER_PIECE": "12"}, {}),
    "pp16": ({"SVDLAS2_SPLIT_PER_PIECE": "16"}, {}),
    "nosplit": ({"SVDLAS2_SPLIT": "0"}, {}), "nowide": ({"SVDLAS2_WIDE_SLABS": "0"}, {}),
    "norelabel": ({"SVDLAS2_RELABEL": "0"}, {}),
    "hot4k": ({}, {"spmv_hot": 4096}), "hot12k": ({}, {"spmv_hot": 12288}), "hot16k": ({}, {"spmv_hot": 16384}),
    "shape16": ({}, {"spmv_shape": 16}),
}
out = open(os.path.join(ROOT, "gpurun_out", "tune_layout.jsonl"), "a")
for name in a.sets.split(","):
    env, knobs = SETS[name]
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        t0 = time.time()
        with make_op() as op:
            up = time.time() - t0
            for k, v in knobs.items():
                op.set(k, v)
            t = op.time_opb(reps=5)
            rec = dict(set=name, workload=a.workload, shard=a.shard, nnz=op.nnz, upload_s=round(up, 2), ms_opb=round(t["ms_per_opb"], 4),
                       ms_b=round(t["ms_spmv_b"], 4), ms_bt=round(t["ms_spmv_bt"], 4), GBs=round(t["bytes_opb"] / t["ms_per_opb"] / 1e6, 1),
                       frac=round(t["bytes_opb"] / t["ms_per_opb"] / 1e6 / 6561.6, 4))
    finally:
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    print("TUNE " + json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n"); out.flush()
