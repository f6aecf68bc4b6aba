"""torchrun target: sharded solve on N GPUs vs the single-GPU solve of rank 0 (and vs the oracle when --oracle).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
       scripts/multi_gpu_check.py [--workload cfg2] [--scale 0.05] [--dims 30]"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svdlibrs_b200 as las2
from svdlibrs_b200 import sharding, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--scale", type=float, default=0.05)
ap.add_argument("--dims", type=int, default=30)
ap.add_argument("--oracle", type=int, default=0)
ap.add_argument("--nside-mode", type=int, default=0, help="2: the replicated N-side vectors are dealt out by rows over the ranks")
ap.add_argument("--seed", type=int, default=12345, help="0: the library draws the seed (rank 0's draw must reach every rank)")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
A, _ = synthetic.make(a.workload, scale=a.scale)
B, transposed = sharding.orient(A)
bounds = sharding.partition_rows_by_nnz(B.indptr, world)
Bl, lo, hi = sharding.local_shard(B, bounds, rank)
comm = las2.Communicator(rank, world, sharding.broadcast_comm_id(dist, rank), device=local)
t0 = time.time()
with las2.ShardedOperator(Bl, B.shape[0], lo, int(B.nnz), comm, transposed, device=local) as op:
    op.set("nside_mode", a.nside_mode)
    r = op.solve(a.dims, random_seed=a.seed)
    t_sharded = time.time() - t0
    seed_used = r.diagnostics.random_seed
    for rep in range(3):
        rr = op.solve(a.dims, random_seed=seed_used)
        if rank == 0:
            print("PROFILE", rep, json.dumps({k: round(rr.profile[k], 4) for k in ("t_total", "t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_download", "t_device")}), flush=True)
        del rr
    tm = op.time_opb(reps=10)
# every rank must hold bit-identical N-side vectors and scalars
sig = torch.tensor(r.s, device="cuda")
ref = sig.clone(); dist.broadcast(ref, 0)
same_s = bool(torch.equal(sig, ref))
nside = r.ut if transposed else r.vt
if a.nside_mode == 2:
    # every rank holds a contiguous block of the rows: re-assemble them (in rank order) on every rank
    blocks = [None] * world
    dist.all_gather_object(blocks, (r.nside_rows[0], np.array(nside)))
    blocks.sort(key=lambda b: b[0])
    assert [b[0] for b in blocks] == [r.d * g // world for g in range(world)] and sum(b[1].shape[0] for b in blocks) == r.d
    nside = np.concatenate([b[1] for b in blocks], axis=0)
    same_v = True
else:
    v0 = torch.tensor(nside, device="cuda"); vr = v0.clone(); dist.broadcast(vr, 0)
    same_v = bool(torch.equal(v0, vr))
# gather the M-side slices on rank 0
mside = r.vt if transposed else r.ut
parts = [None] * world
dist.all_gather_object(parts, (lo, mside))
ok = [None] * world
dist.all_gather_object(ok, (same_s, same_v, r.diagnostics.lanczos_steps, seed_used))
if rank == 0:
    M_full = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])], axis=1)
    single = las2.svd_dim_seed(A, a.dims, seed_used)
    s_m = single.vt if transposed else single.ut
    s_n = single.ut if transposed else single.vt
    out = dict(world=world, workload=a.workload, shape=list(A.shape), nnz=int(A.nnz), transposed=transposed,
               bounds=[int(b) for b in bounds], all_ranks_identical=all(o[0] and o[1] for o in ok),
               steps_sharded=r.diagnostics.lanczos_steps, steps_single=single.diagnostics.lanczos_steps,
               d_sharded=r.d, d_single=single.d,
               accepted_sharded=r.diagnostics.singular_values, accepted_single=single.diagnostics.singular_values,
               max_rel_dsigma=float(np.max(np.abs(r.s - single.s) / single.s)) if r.d == single.d else None,
               min_cos_n=float(min(abs(nside[i] @ s_n[i]) for i in range(min(r.d, single.d)))),
               min_cos_m=float(min(abs(M_full[i] @ s_m[i]) for i in range(min(r.d, single.d)))),
               t_sharded=t_sharded, ms_per_opb=tm["ms_per_opb"], seed_requested=a.seed, seeds_used=[int(o[3]) for o in ok])
    if a.oracle:
        from oracle import oracle as O
        o = O.svd_dim_seed(A, a.dims, seed_used)
        out.update(steps_oracle=o.diagnostics["lanczos_steps"], accepted_oracle=o.diagnostics["singular_values"],
                   max_rel_dsigma_vs_oracle=float(np.max(np.abs(r.s - o.s) / o.s)) if o.d == r.d else None)
    print("MULTI " + json.dumps(out), flush=True)
dist.destroy_process_group()
