"""Run BASELINE.json configs end to end on one GPU and record time-to-SVD + invariants (no oracle at these sizes).
   python scripts/run_configs.py cfg3:1.0 cfg5:1.0 cfg4:0.25   -> gpurun_out/configs_r1.jsonl"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic

import scipy.sparse as _sp
las2.svd_dim_seed(_sp.random(64, 48, density=0.2, format="csr", random_state=1), 4, 1)  # CUDA context + module load, outside every timing
out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs_r1.jsonl"), "a")
for spec in sys.argv[1:]:
    name, scale = spec.split(":")
    scale = float(scale)
    t0 = time.time()
    A, dims = synthetic.make(name, scale=scale)
    if scale < 1.0:
        dims = synthetic.CONFIGS[name]["dims"] if min(A.shape) > synthetic.CONFIGS[name]["dims"] else dims
    gen = time.time() - t0
    rec = dict(config=name, scale=scale, shape=list(A.shape), nnz=int(A.nnz), fmt=A.format, dims=int(dims), gen_s=round(gen, 1))
    try:
        t0 = time.perf_counter()
        r = las2.svd_dim_seed(A, dims, 12345)
        e2e = time.perf_counter() - t0
        pe = dict(r.profile)
        with las2.Operator(A) as op:
            t0 = time.perf_counter(); r2 = op.solve(dims, random_seed=12345); res = time.perf_counter() - t0
            tm = op.time_opb(reps=10)
        p = r2.profile
        g = r.diagnostics
        ortho = float(np.abs(r.vt[:16] @ r.vt[:16].T - np.eye(min(16, r.d))).max()) if r.d else None
        i = 0
        resid = float(np.linalg.norm(A @ r.vt[i] - r.s[i] * r.ut[i]) / r.s[0]) if r.d else None
        rec.update(e2e_s=round(e2e, 3), resident_s=round(res, 3), d=r.d, lanczos_steps=g.lanczos_steps, transposed=g.transposed,
                   accepted=g.singular_values, stabilized=g.ritz_values_stabilized, sigma0=float(r.s[0]) if r.d else None,
                   sigma_last=float(r.s[-1]) if r.d else None, deterministic=bool(r == r2), ortho_defect16=ortho, resid0=resid,
                   ms_per_opb=round(tm["ms_per_opb"], 4), opb_GBs=round(tm["bytes_opb"] / tm["ms_per_opb"] / 1e6, 1),
                   phases={k: round(p[k], 4) for k in ("t_upload", "t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_download")},
                   e2e_phases={k: round(pe[k], 4) for k in ("t_upload", "t_lanczos", "t_ritvec", "t_imtql2", "t_total")},
                   purge_fired=int(p["n_purge_fired"]), purge_vectors=int(p["n_purge_vectors"]), launches=int(p["n_kernel_launches"]))
    except las2.SvdLibError as e:
        rec.update(error=str(e))
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n"); out.flush()
    del A
