"""Row-sharded solve of a BASELINE config where every rank GENERATES ONLY ITS OWN ROW RANGE of the matrix (cfg 4 is 16 GB of
host arrays: too big to build once per rank).  Works for world_size 1 too (python scripts/scale_cfg.py ...) so that the
1-GPU and N-GPU numbers come from the same code path.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
       scripts/scale_cfg.py --workload cfg4 [--scale 1.0] [--solves 2]
Prints one JSON line (rank 0): time-to-SVD (resident, max over ranks), ms per opb, Lanczos steps, phases, invariants."""
import argparse, json, os, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--solves", type=int, default=2); ap.add_argument("--dims", type=int, default=0)
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
import svdlibrs_b200 as las2
from svdlibrs_b200 import sharding, synthetic
dist = None

cfg = synthetic.CONFIGS[a.workload]
assert cfg["kind"] == "powerlaw" and cfg["fmt"] == "csr" and cfg["ncols"] < 1.2 * cfg["nrows"], "row-sharded generator: tall power-law CSR configs"
nrows, ncols, nnz_req = max(8, int(cfg["nrows"] * a.scale)), max(8, int(cfg["ncols"] * a.scale)), max(16, int(cfg["nnz"] * a.scale))
dims = a.dims or cfg["dims"]
# balance by the REQUESTED row lengths (the realised ones differ by the merged duplicate columns, a few per cent)
lens = np.minimum(synthetic._row_lengths(nrows, nnz_req, 0.6, 12345), ncols)
offs = np.concatenate([[0], np.cumsum(lens)])
bounds = sharding.partition_rows_by_nnz(offs, world)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
t0 = time.time()
try:
    cores = len(os.sched_getaffinity(0))
except Exception:
    cores = os.cpu_count() or 1
Bl = synthetic.powerlaw_csr(nrows, ncols, nnz_req, 12345, rows=(lo, hi), workers=max(1, min(16, cores // world)))
gen_s = time.time() - t0
nnz_local = int(Bl.nnz)
import torch  # CUDA starts only now: the generator forked its workers above
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if world > 1:
    t = torch.tensor([nnz_local], device="cuda", dtype=torch.int64); dist.all_reduce(t); nnz_global = int(t.item())
    comm = las2.Communicator(rank, world, sharding.broadcast_comm_id(dist, rank), device=local)
else:
    nnz_global, comm = nnz_local, None

def vmax(x):
    if world == 1: return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

t0 = time.time()
with las2.ShardedOperator(Bl, nrows, lo, nnz_global, comm, False, device=local) as op:
    upload_s = vmax(time.time() - t0)
    op.set("nside_mode", int(os.environ.get("NSIDE_MODE", "2")))  # V is replicated: every rank downloads 1/G of its rows
    del Bl
    recs = []
    for i in range(a.solves):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = op.solve(dims, random_seed=12345)
        torch.cuda.synchronize(); dt = vmax(time.perf_counter() - t0)
        p = r.profile
        recs.append(dict(time_to_svd_s=round(dt, 4), **{k: round(p[k], 4) for k in ("t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2")}))
        last = r
        if i + 1 < a.solves: del r
    tm = op.time_opb(reps=5)
    g = last.diagnostics
    ortho = float(np.abs(last.vt[:8] @ last.vt[:8].T - np.eye(min(8, last.d))).max()) if last.d else None
    rec = dict(workload=a.workload, scale=a.scale, world=world, shape=[nrows, ncols], nnz=nnz_global, dims=dims, gen_s=round(gen_s, 1),
               upload_s=round(upload_s, 3), solves=recs, d=last.d, lanczos_steps=g.lanczos_steps, accepted=g.singular_values,
               sigma0=float(last.s[0]) if last.d else None, sigma_last=float(last.s[-1]) if last.d else None, ortho_defect8=ortho,
               ms_per_opb=round(vmax(tm["ms_per_opb"]), 4), ms_spmv_b=round(tm["ms_spmv_b"], 4), ms_spmv_bt=round(tm["ms_spmv_bt"], 4),
               opb_bytes_local=tm["bytes_opb"], purge_fired=int(last.profile["n_purge_fired"]), purge_vectors=int(last.profile["n_purge_vectors"]),
               local_rows=[lo, hi], local_nnz=nnz_local, ring=bool(op.info("ring")) if world > 1 else None,
               ring_env=os.environ.get("SVDLAS2_RING", ""))
    if rank == 0:
        print("SCALE " + json.dumps(rec), flush=True)
        if a.out:
            with open(a.out, "a") as f: f.write(json.dumps(rec) + "\n")
if world > 1: dist.destroy_process_group()
