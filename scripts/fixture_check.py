"""One full-size GPU solve of a BASELINE config against the CPU oracle's recorded full-size run (tests/golden/oracle_<cfg>.json,
made by scripts/make_oracle_fixtures.py): identical step / stabilised / accepted / returned counts, singular values to 1e-10,
the sketch of U and V (sampled coordinates + projections) inside what |cos| >= 1 - 1e-8 implies, and the size-independent
invariants (sigma descending, V V' = I, A v = sigma u, A' u = sigma v) on the GPU result itself.

    python scripts/fixture_check.py cfg4        -> one line "FIXTURE {json}"
The matrix is generated BEFORE CUDA is initialised (forked generator workers), which is why this is a script of its own and
not a function inside the pytest process."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from make_oracle_fixtures import sketch_plan  # noqa: E402

name = sys.argv[1]
fx = json.load(open(os.path.join(ROOT, "tests", "golden", f"oracle_{name}.json")))

import scipy.sparse as sp  # noqa: E402
from svdlibrs_b200 import synthetic  # noqa: E402

try:
    cores = len(os.sched_getaffinity(0))
except (AttributeError, OSError):
    cores = os.cpu_count() or 1
t0 = time.time()
A, dims = synthetic.make(name, workers=max(1, min(32, cores)))
gen_s = time.time() - t0
mx = fx["matrix"]
assert list(A.shape) == mx["shape"] and int(A.nnz) == mx["nnz"], "the generator no longer reproduces the fixture's matrix"
assert abs(float(A.data.sum()) - mx["values_sum"]) <= 1e-9 * mx["values_sum"]

import svdlibrs_b200 as las2  # noqa: E402

t0 = time.time()
r = las2.svd_dim_seed(A, dims, 12345)
solve_s = time.time() - t0
g, ref = r.diagnostics, fx["diagnostics"]
out = dict(config=name, gen_s=round(gen_s, 1), e2e_s=round(solve_s, 2), d=[r.d, fx["d"]])
for k in ("non_zero", "dimensions", "iterations", "transposed", "lanczos_steps", "ritz_values_stabilized", "significant_values",
          "singular_values", "random_seed"):
    out[k] = [getattr(g, k), ref[k]]
s_ref = np.array(fx["s"])
out["max_rel_dsigma"] = float(np.max(np.abs(r.s - s_ref) / s_ref)) if r.d == len(s_ref) else None
k = min(len(fx["alf_head"]), 30)
out["alf_head_rel"] = float(np.max(np.abs(r.trace["alf"][:k] - np.array(fx["alf_head"][:k])) / np.abs(np.array(fx["alf_head"][:k]))))

# sketch of U and V: with |cos(v, v_ref)| >= 1 - 1e-8 and unit vectors, |v - +-v_ref| <= sqrt(2e-8) = 1.4143e-4, hence the
# same bound on any sub-vector and |(v - v_ref) . w| <= 1.4143e-4 |w| for any direction w
sk = fx["sketch"]
bound = float(np.sqrt(2e-8)) * 1.001
worst = dict(ut_samples=0.0, vt_samples=0.0, ut_proj=0.0, vt_proj=0.0)
if r.d == fx["d"]:
    for tag, M, plan_tag in (("ut", r.ut, 1), ("vt", r.vt, 2)):
        idx, omega = sketch_plan(M.shape[1], plan_tag)
        assert [int(x) for x in idx] == sk[f"{tag}_idx"]
        wn = np.linalg.norm(omega, axis=1)
        samp, proj = M[:, idx], M @ omega.T
        rs, rp = np.array(sk[f"{tag}_samples"]), np.array(sk[f"{tag}_proj"])
        for i in range(r.d):
            sign = 1.0 if np.dot(proj[i], rp[i]) + np.dot(samp[i], rs[i]) >= 0 else -1.0
            worst[f"{tag}_samples"] = max(worst[f"{tag}_samples"], float(np.linalg.norm(sign * samp[i] - rs[i])))
            worst[f"{tag}_proj"] = max(worst[f"{tag}_proj"], float(np.max(np.abs(sign * proj[i] - rp[i]) / wn)))
out["sketch_worst"] = worst
out["sketch_bound"] = bound
# invariants on the GPU result
out["sigma_descending"] = bool(np.all(np.diff(r.s) <= 0))
out["vt_gram_defect"] = float(np.abs(r.vt[:16] @ r.vt[:16].T - np.eye(min(16, r.d))).max())
res_av, res_atu = 0.0, 0.0
for i in sorted({0, 1, r.d // 2, r.d - 1}):
    res_av = max(res_av, float(np.linalg.norm(A @ r.vt[i] - r.s[i] * r.ut[i]) / r.s[0]))
    res_atu = max(res_atu, float(np.linalg.norm(A.T @ r.ut[i] - r.s[i] * r.vt[i]) / r.s[0]))
out["resid_av"], out["resid_atu"] = res_av, res_atu
out["profile"] = {k: round(r.profile[k], 3) for k in ("t_upload", "t_lanczos", "t_ritvec", "t_imtql2", "t_total")}
out["counts_equal"] = all(out[k][0] == out[k][1] for k in ("d", "non_zero", "dimensions", "iterations", "transposed", "lanczos_steps",
                                                             "ritz_values_stabilized", "significant_values", "singular_values"))
print("FIXTURE " + json.dumps(out), flush=True)
