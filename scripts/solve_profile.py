"""One resident solve with the per-phase profile printed (no oracle).  python scripts/solve_profile.py [--workload cfg2] [--scale 1] [--reps 3]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--dims", type=int, default=0)
a = ap.parse_args()
try:
    cores = len(os.sched_getaffinity(0))
except Exception:
    cores = os.cpu_count() or 1
A, dims = synthetic.make(a.workload, scale=a.scale, workers=min(32, cores))
dims = a.dims or dims
def throttle():
    """CFS throttling counters of this container's cpu cgroup (v1 or v2), or None"""
    for path in ("/sys/fs/cgroup/cpu/cpu.stat", "/sys/fs/cgroup/cpu.stat"):
        try:
            d = dict(l.split() for l in open(path).read().strip().splitlines())
            return {k: int(v) for k, v in d.items() if "throttled" in k}
        except Exception:
            pass
    return None
for path in ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpu.max"):
    try:
        print("cgroup", path, open(path).read().strip(), flush=True)
    except Exception:
        pass
with las2.Operator(A) as op:
    r = None
    for i in range(a.reps):
        r = None  # hand the previous result's pinned buffers back to the library's pool first
        th0 = throttle()
        t0 = time.perf_counter(); r = op.solve(dims, random_seed=12345); dt = time.perf_counter() - t0
        th1 = throttle()
        if th0 and th1:
            print("throttle delta", {k: th1[k] - th0[k] for k in th1}, flush=True)
        p = r.profile
        print(json.dumps(dict(wall=round(dt, 4), steps=r.diagnostics.lanczos_steps, d=r.d, n_opb=p["n_opb"],
                              **{k: round(p[k], 4) for k in ("t_total", "t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_download", "t_device")},
                              launches=p["n_kernel_launches"], syncs=p["n_host_syncs"], purge=p["n_purge_fired"])), flush=True)
