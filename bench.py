#!/usr/bin/env python
"""bench.py -- svd_dim_seed on BASELINE.json's configs, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one whole svd_dim_seed(A, dims, seed=12345) of the named synthetic workload.
  metric  opb matvecs/s over the whole solve: (applications of B'(B x) in the solve) / time-to-SVD
          (BASELINE.json: "svd_dim_seed time-to-SVD & opb matvecs/s"); time-to-SVD itself is in `time_to_svd_s`.
  value   operator already resident in HBM (svdlas2_operator_solve), CUDA-event time on the library's stream
  e2e     the reference-facing call svdlas2_csr/_csc with HOST buffers: upload + conversion + solve + D2H of ut/s/vt
  roofline  the dominant kernel (k_spmv_chunk, two launches per opb) timed live with CUDA events
  cpu_baseline  the CPU oracle (oracle/, a restatement of the reference: the Rust crate cannot be built here),
          one thread like the reference, on a bounded sample (same matrix, `iterations` capped)
--impl reference times that oracle alone with the same metric / config.
N > 1: rows of the oriented operator are sharded by nnz, one NCCL all-reduce per opb (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
METRIC = "svd_dim_seed opb matvecs/s (whole solve)"
UNIT = "opb/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self) -> int:
        """index of the next sample: brackets the timed region (the sampler itself starts before the warm-up so
        that spawning nvidia-smi, which briefly blocks CUDA calls, stays out of the timed steps)"""
        return len(self.rows)

    def stop(self, first: int = 0, last: int | None = None) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        rows = self.rows[first:(last if last is not None and last > first else None)] or self.rows
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workload
def make_workload(name: str, scale: float):
    from svdlibrs_b200 import synthetic
    t0 = time.time()
    A, dims = synthetic.make(name, seed=SEED, scale=scale)
    rl = np.diff(A.indptr)
    info = {"workload": f"{name}: {synthetic.CONFIGS[name]['kind']} {A.format.upper()} {A.shape[0]}x{A.shape[1]}, "
                        f"{A.nnz} nnz, svd_dim_seed(dims={dims}, seed={SEED})",
            "name": name, "shape": list(A.shape), "nnz": int(A.nnz), "format": A.format, "dims": int(dims),
            "seed": SEED, "matrix_seed": SEED, "scale": scale,
            "lane_len_max": int(rl.max()), "lane_len_mean": float(rl.mean()), "gen_s": round(time.time() - t0, 1)}
    return A, dims, info


def oracle_sample(A, dims: int, sample_iterations: int):
    """The CPU oracle on a bounded sample: same matrix, Lanczos steps capped by `iterations`."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    r = O.svd_las2(A, dims, sample_iterations, (-1.0e-30, 1.0e-30), 1.0e-6, SEED)
    dt = time.perf_counter() - t0
    return r, dt


def pick_sample_iterations(nnz: int, dims: int, seconds: float = 10.0) -> int:
    # ~4 ns per nnz per opb on one host core: aim at `seconds` of opb work, at least dims steps
    per_opb = max(1e-4, 4.5e-9 * nnz)
    return int(max(dims, min(4 * dims, seconds / per_opb)))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    A, dims, info = make_workload(args.workload, args.scale)
    # the whole --warmup W --steps K run is meant to end within a few minutes: ~150 s of CPU work shared by the steps
    per_step_s = min(10.0, max(2.0, 150.0 / max(1, args.warmup + args.steps)))
    it = args.sample_iterations or pick_sample_iterations(A.nnz, dims, per_step_s)
    it = min(it, min(A.shape))
    times, nopb = [], 0
    # a single-threaded CPU run has nothing to warm up beyond the first touch of the matrix: one untimed step at most, so
    # that K timed steps of ~20 s (the smallest sample svdLAS2 allows at dims=100: iterations >= dimensions) stay within minutes
    warm = min(args.warmup, 1)
    for i in range(warm + args.steps):
        r, dt = oracle_sample(A, dims, it)
        nopb = r.trace["n_opb"]
        if i >= warm:
            times.append(dt)
        log(f"[reference] step {i}: {dt:.2f} s, {nopb} opb, lanczos_steps {r.diagnostics['lanczos_steps']}")
    total = sum(times)
    value = nopb * len(times) / total
    sample = (f"oracle svdLAS2 on the same matrix with iterations={it} (of {info['shape']} -> "
              f"{nopb} opb + ritvec per step), 1 thread like the reference")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "warmup_run": warm, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": info,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def pinned_like(a: np.ndarray, torch):
    """Copy a numpy array into pinned host memory and return a numpy view of it (inputs of the e2e call)."""
    a = np.ascontiguousarray(a)
    raw = a.view(np.int64) if a.dtype == np.uint64 else a  # torch has no pinned uint64 path on every build
    t = torch.from_numpy(raw).pin_memory()
    return t.numpy().view(a.dtype), t


def run_ours(args):
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries the one JSON line
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: svdlibrs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import svdlibrs_b200 as las2
    from svdlibrs_b200 import sharding

    A, dims, info = make_workload(args.workload, args.scale)
    endi, kappa = (-1.0e-30, 1.0e-30), 1.0e-6
    info["l2_policy"] = "inputs larger than L2 (operator >= 1 GB vs 126 MB L2); no flush between steps" \
        if 24 * A.nnz > 4 * 126e6 else "operator fits L2: L2 flushed (256 MB memset) before every timed opb of the roofline leg"
    info["parallelism"] = "single GPU" if world == 1 else f"rows of B sharded by nnz over {world} GPUs, 1 NCCL all-reduce of N f64 per opb"

    # ---------------- host inputs in pinned memory (the arrays a nalgebra-sparse matrix would hand over, u64 indices)
    if world == 1:
        hp, _k1 = pinned_like(A.indptr.astype(np.uint64), torch)
        hi, _k2 = pinned_like(A.indices.astype(np.uint64), torch)
        hv, _k3 = pinned_like(A.data.astype(np.float64), torch)
        A_pinned = las2.RawMatrix(A.format, A.shape, hp, hi, hv)
        h2d_step = hp.nbytes + hi.nbytes + hv.nbytes

        def make_resident():
            return las2.Operator(A_pinned, device=local)

        def e2e_call():
            return las2.svdLAS2(A_pinned, dims, 0, endi, kappa, SEED)
    else:
        B, transposed = sharding.orient(A)
        bounds = sharding.partition_rows_by_nnz(B.indptr, world)
        Bl, lo, hi_ = sharding.local_shard(B, bounds, rank)
        hp, _k1 = pinned_like(Bl.indptr.astype(np.uint64), torch)
        hi, _k2 = pinned_like(Bl.indices.astype(np.uint64), torch)
        hv, _k3 = pinned_like(Bl.data.astype(np.float64), torch)
        Bl_pinned = las2.RawMatrix("csr", Bl.shape, hp, hi, hv)
        h2d_step = hp.nbytes + hi.nbytes + hv.nbytes
        comm = las2.Communicator(rank, world, sharding.broadcast_comm_id(dist, rank), device=local)

        def make_resident():
            return las2.ShardedOperator(Bl_pinned, B.shape[0], lo, int(B.nnz), comm, transposed, device=local)

        def e2e_call():
            with make_resident() as op_:
                return op_.solve(dims, 0, endi, kappa, SEED)

    # ---------------- resident leg: value
    op = make_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    res = None
    for _ in range(args.warmup):
        res = None
        res = op.solve(dims, 0, endi, kappa, SEED)
    barrier()
    mark0 = sampler.mark()
    t_dev, launches, wall0 = 0.0, 0, time.perf_counter()
    for _ in range(args.steps):
        res = None  # hand the previous result's pinned buffers back to the library's pool before the next call
        res = op.solve(dims, 0, endi, kappa, SEED)
        t_dev += res.profile["t_device"]
        launches += res.profile["n_kernel_launches"]
        if rank == 0:
            log("[resident] " + json.dumps({k: round(res.profile[k], 4) for k in
                                            ("t_total", "t_lanczos", "t_ritvec", "t_imtql2", "t_device")}))
    barrier()
    wall_resident = time.perf_counter() - wall0
    t_dev = max_over_ranks(t_dev)
    # numerator: the opb applications the REFERENCE algorithm performs for this solve (one per Lanczos
    # step, one in startv, one per restart, one per returned triplet) -- the same for both arms because step counts
    # are identical; the library itself needs fewer (sigma = |B v| saves the tail's second SpMV), see solve.n_opb_run
    n_opb = res.diagnostics.lanczos_steps + 1 + res.d + int(res.profile["n_restarts"])
    value = n_opb * args.steps / t_dev

    # ---------------- roofline leg: the dominant kernel, timed live with CUDA events on the library's stream
    flush = 24 * A.nnz <= 4 * 126e6
    tm = op.time_opb(reps=20, flush_l2=flush)
    barrier()
    clocks = sampler.stop(mark0, sampler.mark() + 1) if rank == 0 else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    ach = tm["bytes_opb"] / 2 / (tm["ms_per_opb"] / 2 * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("workload") == info["name"] and t.get("scale") == info["scale"]:
            traffic = t.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_spmv_chunk (2 launches per opb: B then B')", "achieved": ach,
                "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": tm["bytes_opb"] / 2, "ms_per_launch": tm["ms_per_opb"] / 2,
                "spmv_b": {"GB/s": tm["bytes_spmv_b"] / tm["ms_spmv_b"] / 1e6, "ms": tm["ms_spmv_b"]},
                "spmv_bt": {"GB/s": tm["bytes_spmv_bt"] / tm["ms_spmv_bt"] / 1e6, "ms": tm["ms_spmv_bt"]},
                "frac_of_nominal_8TBs": ach / 8000.0}
    op.close()

    # ---------------- e2e leg: host buffers in, host SvdRec out, every step
    r2 = None
    for _ in range(args.warmup):
        r2 = e2e_call()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        r2 = None
        r2 = e2e_call()
        d2h = r2.profile["d2h_bytes"]
        if rank == 0:
            log("[e2e] " + json.dumps({k: round(r2.profile[k], 4) for k in
                                       ("t_total", "t_upload", "t_lanczos", "t_ritvec", "t_imtql2", "t_device")}))
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_opb * args.steps / t_e2e
    same = (r2.d == res.d and np.array_equal(r2.s, res.s) and r2.diagnostics == res.diagnostics)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h), "time_to_svd_s": t_e2e / args.steps},
            "time_to_svd_s": {"resident": t_dev / args.steps, "resident_wall": wall_resident / args.steps,
                              "e2e": t_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "solve": {"lanczos_steps": res.diagnostics.lanczos_steps, "n_opb": int(n_opb), "n_opb_run": int(res.profile["n_opb"]),
                      "n_spmv_run": int(res.profile["n_spmv"]),
                      "ritz_values_stabilized": res.diagnostics.ritz_values_stabilized, "d": res.d,
                      "accepted": res.diagnostics.singular_values, "sigma_max": float(res.s[0]) if res.d else None,
                      "purge_fired": int(res.profile["n_purge_fired"]), "purge_vectors": int(res.profile["n_purge_vectors"]),
                      "host_syncs": int(res.profile["n_host_syncs"]),
                      "phases_s": {k: res.profile[k] for k in ("t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_download")},
                      "e2e_phases_s": {k: r2.profile[k] for k in ("t_upload", "t_lanczos", "t_ritvec", "t_download", "t_total")},
                      "e2e_equals_resident": bool(same)}}

    # ---------------- cpu_baseline: the oracle on this box's host cores, bounded sample (rank 0, N == 1 only)
    if world == 1 and not args.no_cpu_baseline:
        it = min(args.sample_iterations or pick_sample_iterations(A.nnz, dims), min(A.shape))
        r, dt = oracle_sample(A, dims, it)
        line["cpu_baseline"] = {"value": r.trace["n_opb"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"oracle svdLAS2, same matrix, iterations={it}: {r.trace['n_opb']} opb in {dt:.1f} s "
                                          f"(opb {r.trace['t_opb']:.1f} s, purge {r.trace['t_purge']:.1f} s), 1 thread like the reference",
                                "host_cpus": os.cpu_count()}
    else:
        line["cpu_baseline"] = None
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", help="cfg1..cfg5 of BASELINE.json (default cfg2 = configs[1])")
    ap.add_argument("--scale", type=float, default=1.0, help="scale every dimension (tests only; 1.0 = the named config)")
    ap.add_argument("--sample-iterations", type=int, default=0, help="Lanczos-step cap of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: fewer than 3 warm-up steps requested; the timing rules ask for >= 3")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
