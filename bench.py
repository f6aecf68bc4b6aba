#!/usr/bin/env python
"""bench.py -- svd_dim_seed on BASELINE.json's configs, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one whole svd_dim_seed(A, dims, seed=12345) of the named synthetic workload.  The default workload at EVERY N
is cfg4 = BASELINE.json configs[3], the configuration both north-star targets are quoted on (power-law CSR 10M x 2M,
1e9 requested nnz, dims 200); it fits one B200, so N = 1 runs the whole matrix and N > 1 shards its rows (strong scaling).
  metric  opb matvecs/s over the whole solve: (applications of B'(B x) in the solve) / time-to-SVD
          (BASELINE.json: "svd_dim_seed time-to-SVD & opb matvecs/s"); time-to-SVD itself is in `time_to_svd_s`.
  value   operator already resident in HBM (svdlas2_operator_solve), CUDA-event time on the library's stream
  e2e     the reference-facing call svdlas2_csr/_csc with pinned HOST buffers: upload + conversion + solve + D2H of ut/s/vt
  roofline  the dominant kernel (k_spmv_chunk, two launches per opb) timed live with CUDA events
  parity  the solve's step / accepted counts and singular values against the CPU oracle's full-size run of the same matrix and
          seed (tests/golden/oracle_<cfg>.json, produced by scripts/make_oracle_fixtures.py) -- the same at every N
  cpu_baseline  the CPU oracle (oracle/, a restatement of the reference: the Rust crate cannot be built here), one thread
          like the reference, on a BOUNDED sample: a few svd_opb, one purge sweep, the per-step BLAS-1 chain, contraction
          terms and imtql2 at the final js are timed and extrapolated with the counts of the full solve
--impl reference times that oracle alone with the same metric / config (rank 0 only; the other ranks exit 0).
N > 1: every rank generates and uploads ONLY its own row range of the oriented operator (balanced by nnz), one all-reduce
of N doubles per opb; `extra` carries the cfg2 line of earlier rounds when --with-cfg2 is given.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
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
ENDI, KAPPA = (-1.0e-30, 1.0e-30), 1.0e-6


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


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
class Workload:
    """Host arrays of the workload as a nalgebra-sparse matrix would hand them over (u64 indices).  world == 1: the whole
    matrix in its own format; world > 1: this rank's row range [lo, hi) of the oriented operator B as CSR."""
    pass


def make_workload(name: str, scale: float, rank: int = 0, world: int = 1, workers: int = 1) -> Workload:
    from svdlibrs_b200 import sharding, synthetic
    cfg = synthetic.CONFIGS[name]
    nrows, ncols, nnz_req, dims = synthetic.config_dims(name, scale)
    W = Workload()
    W.name, W.scale, W.world = name, scale, world
    t0 = time.time()
    if cfg["kind"] == "powerlaw":
        # the generator builds the TALL matrix T row by row; cfg 3 hands its arrays over as CSC of the wide T'
        t_rows, t_cols = (ncols, nrows) if cfg["fmt"] == "csc" else (nrows, ncols)
        W.transposed = cfg["fmt"] == "csc"  # B = A' = T (cols(A) >= 1.2 rows(A), reference src/lib.rs:606)
        assert W.transposed == (float(ncols) >= 1.2 * float(nrows))
        lo, hi = 0, t_rows
        if world > 1:
            # balance by the REQUESTED row lengths: no rank has to generate the whole matrix
            lens = np.minimum(synthetic._row_lengths(t_rows, nnz_req, 0.6, SEED), t_cols)
            bounds = sharding.partition_rows_by_nnz(np.concatenate([[0], np.cumsum(lens)]), world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        indptr, idx, val = synthetic.powerlaw_arrays(t_rows, t_cols, nnz_req, SEED, rows=(lo, hi), workers=workers)
        W.fmt = "csr" if world > 1 else cfg["fmt"]
        W.shape = (nrows, ncols)               # shape of A
        W.local_shape = (hi - lo, t_cols)      # rows of B held here
        W.m_global, W.row_begin = t_rows, lo
    else:
        A, _ = synthetic.make(name, seed=SEED, scale=scale)
        B, W.transposed = sharding.orient(A)
        m_all = B.shape[0]
        lo, hi = 0, m_all
        if world > 1:
            bounds = sharding.partition_rows_by_nnz(B.indptr, world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            B = B[lo:hi]
            src, W.fmt = B, "csr"
        else:
            src, W.fmt = A, A.format
        indptr, idx, val = src.indptr, src.indices, src.data
        W.shape = A.shape
        W.local_shape = (hi - lo, B.shape[1])
        W.m_global, W.row_begin = m_all, lo
    W.indptr = np.ascontiguousarray(indptr).astype(np.uint64)
    idx = np.ascontiguousarray(idx)
    W.indices = idx.view(np.uint64) if idx.dtype == np.int64 else idx.astype(np.uint64)
    W.data = np.ascontiguousarray(val, dtype=np.float64)
    W.dims = int(min(dims, min(W.shape)))
    W.local_nnz = int(W.data.size)
    W.nnz = W.local_nnz  # completed over the ranks by the caller when world > 1
    lane = np.diff(W.indptr.astype(np.int64))
    W.lane_len_max, W.lane_len_sum, W.lanes = int(lane.max()) if lane.size else 0, int(lane.sum()), int(lane.size)
    W.gen_s = time.time() - t0
    W.kind = cfg["kind"]
    return W


def workload_info(W: Workload) -> dict:
    return {"workload": f"{W.name}: {W.kind} {('CSC' if W.transposed and W.kind == 'powerlaw' else 'CSR')} {W.shape[0]}x{W.shape[1]}, "
                        f"{W.nnz} nnz, svd_dim_seed(dims={W.dims}, seed={SEED})",
            "name": W.name, "shape": list(W.shape), "nnz": int(W.nnz), "format": "csc" if (W.transposed and W.kind == "powerlaw") else "csr",
            "dims": W.dims, "seed": SEED, "matrix_seed": SEED, "scale": W.scale, "lane_len_max": W.lane_len_max,
            "gen_s": round(W.gen_s, 1), "generator_workers": getattr(W, "workers", 1)}


def fixture_for(name: str, scale: float):
    p = os.path.join(ROOT, "tests", "golden", f"oracle_{name}.json")
    if scale == 1.0 and os.path.exists(p):
        with open(p) as f:
            return json.load(f), os.path.relpath(p, ROOT)
    return None, None


# ------------------------------------------------------------------------------------------------ CPU sample (oracle)
_IMTQL2_CACHE = {}


def imtql2_seconds(js: int) -> float:
    """The oracle's imtql2 (O(js^3), reference :1576-1693) on a js x js tridiagonal with the spread of a Lanczos T; measured
    once per js and reused by the later steps of a run."""
    if js not in _IMTQL2_CACHE:
        from oracle import oracle as O
        rng = np.random.default_rng(js)
        d = np.sort(rng.random(js) ** 4)[::-1] * 1e4 + 1.0
        e = np.concatenate([[0.0], rng.random(js - 1) * 0.3 * np.sqrt(d[1:] * d[:-1]) / js ** 0.5])
        t0 = time.perf_counter()
        rc, _, _ = O.imtql2(d, e)
        _IMTQL2_CACHE[js] = time.perf_counter() - t0
        if rc != 0:
            log(f"[cpu sample] imtql2 on the synthetic T returned {rc}")
    return _IMTQL2_CACHE[js]


def cpu_sample(W: Workload, counts: dict, n_opb_timed: int, nvec: int = 8) -> dict:
    """Bounded sample of the reference's CPU path on this workload: `n_opb_timed` real svd_opb applications on
    the real matrix, one purge sweep over `nvec` stored vectors, one BLAS-1 chain of a Lanczos step, `nvec` contraction
    terms, imtql2 at the final js; extrapolated to the whole solve with `counts` (from the oracle's own full run when a
    fixture exists, else from the GPU solve, whose counts are identical).  One thread, like the reference."""
    from oracle import oracle as O
    fmt = {"csr": O.FMT_CSR, "csc": O.FMT_CSC}[W.fmt]
    nr, nc = W.shape
    transposed = W.transposed
    n = nr if transposed else nc
    x = O.start_vector(SEED, n)
    t0 = time.perf_counter()
    for _ in range(n_opb_timed):
        y = O.opb_raw(fmt, nr, nc, W.indptr, W.indices, W.data, x, transposed)
        x = y / np.linalg.norm(y)
    t_opb = (time.perf_counter() - t0) / n_opb_timed
    u = O.unit_costs(n, nvec)
    js = counts["lanczos_steps"]
    t_q = imtql2_seconds(js)
    parts = {"opb": counts["n_opb"] * t_opb, "opa": counts["n_opa"] * t_opb / 2.0,
             "purge": counts["n_purge_vecs"] * u["purge_vec_s"], "blas1": js * u["step_blas1_s"],
             "imtql2": t_q, "contraction": counts["accepted"] * js * u["contract_term_s"]}
    total = sum(parts.values())
    return {"seconds_extrapolated": total, "opb_per_s": counts["n_opb_reference"] / total, "parts_s": {k: round(v, 3) for k, v in parts.items()},
            "measured": {"opb_s": t_opb, "n_opb_timed": n_opb_timed, **u, "imtql2_s": t_q, "js": js}}


def reference_counts(fx, gpu_res=None) -> dict:
    """Counts of the full solve that the extrapolation multiplies the unit costs with."""
    if fx is not None:
        g, c = fx["diagnostics"], fx["counters"]
        steps, d, acc = g["lanczos_steps"], fx["d"], g["singular_values"]
        return {"lanczos_steps": steps, "n_opb": c["n_opb"], "n_opa": c["n_opa"], "n_purge_vecs": c["n_purge_vecs"],
                "accepted": acc, "d": d, "n_opb_reference": steps + 1 + d + c["n_restarts"], "source": "oracle full-size run (fixture)"}
    g, p = gpu_res.diagnostics, gpu_res.profile
    steps, d = g.lanczos_steps, gpu_res.d
    return {"lanczos_steps": steps, "n_opb": steps + 1 + d + int(p["n_restarts"]), "n_opa": d,
            "n_purge_vecs": int(p["n_purge_vectors"]), "accepted": g.singular_values, "d": d,
            "n_opb_reference": steps + 1 + d + int(p["n_restarts"]), "source": "counts of the GPU solve (identical by parity)"}


def sample_text(cs: dict, counts: dict) -> str:
    m = cs["measured"]
    return (f"extrapolated: {m['n_opb_timed']} oracle svd_opb on the same matrix ({m['opb_s']:.2f} s each), one purge sweep "
            f"({m['purge_vec_s'] * 1e3:.1f} ms per stored vector), one step's BLAS-1 chain ({m['step_blas1_s'] * 1e3:.1f} ms), "
            f"contraction terms ({m['contract_term_s'] * 1e3:.2f} ms), imtql2 at js={m['js']} ({m['imtql2_s']:.2f} s); multiplied by the "
            f"counts of the full solve ({counts['n_opb']} opb, {counts['n_opa']} opa, {counts['n_purge_vecs']} purge visits, "
            f"{counts['accepted']} accepted; {counts['source']}) -> {cs['seconds_extrapolated']:.0f} s; 1 thread like the reference")


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    workers = max(1, min(32, host_cores()))
    W = make_workload(args.workload, args.scale, 0, 1, workers)
    W.workers = workers
    info = workload_info(W)
    fx, fx_path = fixture_for(W.name, W.scale)
    steps_total = max(1, args.steps)
    warm = min(args.warmup, 1)  # a single-threaded CPU run has nothing to warm up beyond the first touch of the matrix
    if fx is None:
        return run_reference_truncated(args, W, info, warm)
    counts = reference_counts(fx)
    # size the sample: the whole run is meant to end within a few minutes (~180 s of CPU work over all steps), and at
    # least 10 svd_opb are timed over the run
    t_probe = max(1e-3, 4.5e-9 * W.nnz)
    per_step = max(1, min(16, int(180.0 / (steps_total + warm) / t_probe / 1.3)))
    per_step = max(per_step, int(math.ceil(10.0 / steps_total)))
    times, vals, last = [], [], None
    for i in range(warm + steps_total):
        t0 = time.perf_counter()
        cs = cpu_sample(W, counts, per_step)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt); vals.append(cs["opb_per_s"]); last = cs
        log(f"[reference] step {i}: sample took {dt:.1f} s -> extrapolated solve {cs['seconds_extrapolated']:.0f} s, "
            f"{cs['opb_per_s']:.4f} opb/s {cs['parts_s']}")
    value = len(vals) / sum(1.0 / v for v in vals)  # steps of equal work: total opb / total (extrapolated) time
    measured_full = fx.get("cpu_seconds", {})
    sample = {"kind": "extrapolated", "text": sample_text(last, counts), "opb_timed_per_step": per_step,
              "opb_timed_total": per_step * len(vals), "counts": counts, "parts_s": last["parts_s"], "fixture": fx_path,
              "full_run_on_build_box_s": measured_full.get("total"), "full_run_host": measured_full.get("host"),
              "sample_seconds_per_step": sum(times) / len(times)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "warmup_run": warm, "ms_per_step": 1e3 * counts["n_opb_reference"] / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": info, "time_to_svd_s": {"extrapolated": counts["n_opb_reference"] / value},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample["text"], "sample_detail": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_truncated(args, W, info, warm):
    """No full-size oracle fixture for this workload / scale: run the real oracle svdLAS2 with `iterations` capped."""
    from oracle import oracle as O
    import scipy.sparse as sp
    ctor = sp.csc_matrix if W.fmt == "csc" else sp.csr_matrix
    A = ctor((W.data, W.indices.view(np.int64), W.indptr.view(np.int64)), shape=W.shape)
    per_step_s = min(10.0, max(2.0, 150.0 / max(1, warm + args.steps)))
    it = int(max(W.dims, min(4 * W.dims, per_step_s / max(1e-4, 4.5e-9 * W.nnz))))
    it = min(it, min(W.shape))
    times, nopb = [], 0
    for i in range(warm + args.steps):
        t0 = time.perf_counter()
        r = O.svd_las2(A, W.dims, it, ENDI, KAPPA, SEED)
        dt = time.perf_counter() - t0
        nopb = r.trace["n_opb"]
        if i >= warm:
            times.append(dt)
        log(f"[reference] step {i}: {dt:.2f} s, {nopb} opb, lanczos_steps {r.diagnostics['lanczos_steps']}")
    total = sum(times)
    value = nopb * len(times) / total
    text = (f"truncated: oracle svdLAS2 on the same matrix with iterations={it} ({nopb} opb + ritvec per step), 1 thread like the reference")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "warmup_run": warm, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": info,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": text,
                             "sample_detail": {"kind": "truncated", "iterations": it}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def pin_in_place(arrays, torch):
    """cudaHostRegister the generator's arrays where they lie (no second copy of the 15 GB of cfg 4): these are the pinned
    HOST buffers of the e2e call.  Returns the registered pointers (unregistered at exit)."""
    rt = torch.cuda.cudart()
    done = []
    for a in arrays:
        if a.nbytes == 0:
            continue
        rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        if int(rc) != 0:
            log(f"[bench] cudaHostRegister of {a.nbytes >> 20} MiB failed ({rc}); the e2e leg reads pageable memory")
        else:
            done.append(a.ctypes.data)
    return done


def parity_block(res, fx, fx_path):
    if fx is None:
        return {"checked": False, "why": "no full-size oracle fixture for this workload / scale"}
    g, ref = res.diagnostics, fx["diagnostics"]
    s_ref = np.array(fx["s"])
    out = {"checked": True, "oracle_fixture": fx_path,
           "lanczos_steps": [g.lanczos_steps, ref["lanczos_steps"]],
           "ritz_values_stabilized": [g.ritz_values_stabilized, ref["ritz_values_stabilized"]],
           "significant_values": [g.significant_values, ref["significant_values"]],
           "singular_values": [g.singular_values, ref["singular_values"]],
           "non_zero": [g.non_zero, ref["non_zero"]]}
    counts_ok = all(a == b for a, b in (v for k, v in out.items() if isinstance(v, list)))
    rel = float(np.max(np.abs(res.s - s_ref) / np.abs(s_ref))) if res.d == len(s_ref) and res.d else float("inf")
    out["max_rel_dsigma"] = rel
    out["ok"] = bool(counts_ok and rel <= 1e-10)
    return out


def extra_workload(name: str, local: int, las2, peak: float, steps: int = 5, warmup: int = 3) -> dict:
    """Resident-leg numbers of another BASELINE config (generated single-threaded: CUDA is already up, no forking here)."""
    W = make_workload(name, 1.0, 0, 1, 1)
    A = las2.RawMatrix(W.fmt, W.shape, W.indptr, W.indices, W.data)
    fx, fx_path = fixture_for(name, 1.0)
    with las2.Operator(A, device=local) as op:
        res = None
        for _ in range(warmup):
            res = None
            res = op.solve(W.dims, 0, ENDI, KAPPA, SEED)
        t = 0.0
        for _ in range(steps):
            res = None
            res = op.solve(W.dims, 0, ENDI, KAPPA, SEED)
            t += res.profile["t_device"]
        tm = op.time_opb(reps=20, flush_l2=24 * W.nnz <= 4 * 126e6)
        n_opb = res.diagnostics.lanczos_steps + 1 + res.d + int(res.profile["n_restarts"])
        ach = tm["bytes_opb"] / tm["ms_per_opb"] / 1e6
        out = {"workload": workload_info(W)["workload"], "time_to_svd_s": t / steps, "value": n_opb * steps / t, "unit": UNIT,
               "steps": steps, "warmup": warmup, "lanczos_steps": res.diagnostics.lanczos_steps,
               "roofline": {"achieved": ach, "peak": peak, "frac": ach / peak, "ms_per_opb": tm["ms_per_opb"]},
               "parity": parity_block(res, fx, fx_path)}
        res = None
    return out


def run_ours(args):
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # ---------------- the workload is generated BEFORE CUDA is initialised (forked generator workers)
    workers = max(1, min(32, host_cores() // max(1, env_int("LOCAL_WORLD_SIZE", world))))
    W = make_workload(args.workload, args.scale, rank, world, workers)
    W.workers = workers

    import torch
    import torch.distributed as dist
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

    def reduce_over_ranks(x: float, op="max") -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    import svdlibrs_b200 as las2
    from svdlibrs_b200 import sharding

    W.nnz = int(round(reduce_over_ranks(float(W.local_nnz), "sum")))
    W.lane_len_max = int(reduce_over_ranks(float(W.lane_len_max)))
    W.gen_s = reduce_over_ranks(W.gen_s)
    info = workload_info(W)
    dims = W.dims
    info["l2_policy"] = "inputs larger than L2 (operator >= 1 GB per GPU vs 126 MB L2); no flush between steps" \
        if 24 * W.local_nnz > 4 * 126e6 else "operator fits L2: L2 flushed (256 MB memset) before every timed opb of the roofline leg"
    info["parallelism"] = "single GPU" if world == 1 else \
        f"rows of B sharded by nnz over {world} GPUs (each rank generates and uploads only its rows), 1 all-reduce of N f64 per opb"
    fx, fx_path = fixture_for(W.name, W.scale)

    # ---------------- host inputs: pinned in place (the arrays a nalgebra-sparse matrix would hand over, u64 indices)
    pinned = pin_in_place([W.indptr, W.indices, W.data], torch)
    h2d_step = W.indptr.nbytes + W.indices.nbytes + W.data.nbytes
    if world == 1:
        A_host = las2.RawMatrix(W.fmt, W.shape, W.indptr, W.indices, W.data)
        comm = None

        def make_resident():
            return las2.Operator(A_host, device=local)

        def e2e_call():
            return las2.svdLAS2(A_host, dims, 0, ENDI, KAPPA, SEED)
    else:
        B_host = las2.RawMatrix("csr", W.local_shape, W.indptr, W.indices, W.data)
        comm = las2.Communicator(rank, world, sharding.broadcast_comm_id(dist, rank), device=local)

        def make_resident():
            op_ = las2.ShardedOperator(B_host, W.m_global, W.row_begin, W.nnz, comm, W.transposed, device=local)
            op_.set("nside_mode", 2)  # V is replicated: every rank brings 1/G of its rows to the host (U comes out row-sliced per rank)
            return op_

        e2e_parts = {"create": 0.0, "solve": 0.0, "destroy": 0.0}

        def e2e_call():
            t0 = time.perf_counter()
            op_ = make_resident()
            t1 = time.perf_counter()
            try:
                r = op_.solve(dims, 0, ENDI, KAPPA, SEED)
            finally:
                t2 = time.perf_counter()
                op_.close()
            t3 = time.perf_counter()
            e2e_parts["create"], e2e_parts["solve"], e2e_parts["destroy"] = t1 - t0, t2 - t1, t3 - t2
            return r

    # ---------------- resident leg: value
    t0 = time.perf_counter()
    op = make_resident()
    upload_first_s = reduce_over_ranks(time.perf_counter() - t0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    res = None
    for _ in range(args.warmup):
        res = None
        res = op.solve(dims, 0, ENDI, KAPPA, SEED)
    barrier()
    mark0 = sampler.mark()
    t_dev, launches, wall0 = 0.0, 0, time.perf_counter()
    for _ in range(args.steps):
        res = None  # hand the previous result's pinned buffers back to the library's pool before the next call
        res = op.solve(dims, 0, ENDI, KAPPA, SEED)
        t_dev += res.profile["t_device"]
        launches += res.profile["n_kernel_launches"]
        if rank == 0:
            log("[resident] " + json.dumps({k: round(res.profile[k], 4) for k in
                                            ("t_total", "t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_device")}))
    barrier()
    wall_resident = time.perf_counter() - wall0
    t_dev = reduce_over_ranks(t_dev)
    # numerator: the opb applications the REFERENCE algorithm performs for this solve (one per Lanczos
    # step, one in startv, one per restart, one per returned triplet) -- the same for both arms because step counts
    # are identical; the library itself needs fewer (sigma = |B v| saves the tail's second SpMV), see solve.n_opb_run
    n_opb = res.diagnostics.lanczos_steps + 1 + res.d + int(res.profile["n_restarts"])
    value = n_opb * args.steps / t_dev
    parity = parity_block(res, fx, fx_path)
    if parity.get("checked") and not parity["ok"]:
        log("[bench] PARITY MISMATCH against the oracle fixture: " + json.dumps(parity))

    # ---------------- roofline leg: the dominant kernel, timed live with CUDA events on the library's stream
    flush = 24 * W.local_nnz <= 4 * 126e6
    tm = op.time_opb(reps=20 if W.local_nnz < 3e8 else 10, flush_l2=flush)
    exchange = None
    if world > 1:
        ring = bool(op.info("ring"))
        exchange = {"kind": "peer-memory kernel over NVLink, fused with r -= rnm q and the alf partials (rank-ordered sum)" if ring
                    else "ncclAllReduce + separate update kernel",
                    "ring": ring, "two_shot": bool(op.info("ring_two_shot_calls") > 0),
                    "nvlink_bytes_per_rank_per_opb": op.info("ring_bytes_per_call") if ring else None,
                    "calls": int(op.info("ring_calls"))}
    relabelled = bool(op.info("relabelled"))
    hot_fraction = op.info("hot_fraction")
    barrier()
    clocks = sampler.stop(mark0, sampler.mark() + 1) if rank == 0 else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    ms_opb = reduce_over_ranks(tm["ms_per_opb"])
    ach = tm["bytes_opb"] / 2 / (tm["ms_per_opb"] / 2 * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        ent = t.get(W.name) if isinstance(t.get(W.name), dict) else (t if t.get("workload") == W.name else None)
        if ent and ent.get("scale", 1.0) == W.scale and ent.get("n_gpus", 1) == world:
            traffic = ent.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_spmv_chunk (2 launches per opb: B then B')" + ("" if world == 1 else " on rank 0's shard; B' time includes the all-reduce"),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": tm["bytes_opb"] / 2, "ms_per_launch": tm["ms_per_opb"] / 2,
                "ms_per_opb_max_over_ranks": ms_opb,
                "spmv_b": {"GB/s": tm["bytes_spmv_b"] / tm["ms_spmv_b"] / 1e6, "ms": tm["ms_spmv_b"]},
                "spmv_bt": {"GB/s": tm["bytes_spmv_bt"] / tm["ms_spmv_bt"] / 1e6, "ms": tm["ms_spmv_bt"]},
                "frac_of_nominal_8TBs": ach / 8000.0,
                # what stands between `frac` and 1 (evidence under profiles/): the kernel is HBM-bound by contract, but half of
                # this matrix's gathers are random on both sides and stay on the L1 path
                "measured_limiter": "L1TEX tag stage: ~0.95 uncoalesced 8-byte gathers per clock per SM (2.6 from shared memory); "
                                    "0.70 of the HBM peak needs 1.5 per clock at the sustained clock (profiles/r2g_ncu_summary.md, "
                                    "r1_gather_microbench.md, r2u_tma_gather_microbench.md)"}
    op.close()
    # keep only what the line needs of the resident result: its 19 GB of pinned ut / vt (cfg 4) go back to the library's host
    # pool NOW, otherwise the e2e leg's results do not fit the pool's cache next to them and every e2e call would create and
    # register its buffers afresh (hundreds of ms each, and cudaHostRegister stalls the device while it runs)
    class _Kept:
        pass
    kept = _Kept()
    kept.d, kept.s, kept.diagnostics, kept.profile = res.d, res.s.copy(), res.diagnostics, dict(res.profile)
    res = kept

    # ---------------- e2e leg: host buffers in, host SvdRec out, every step.  Same K / W unless that would not fit the
    # driver's per-run limit: then fewer steps (never below 3), stated in the line.
    budget = float(os.environ.get("BENCH_E2E_BUDGET_S", "300"))
    est = 1.25 * (t_dev / args.steps) + upload_first_s
    e2e_steps = args.steps if est * (args.steps + args.warmup) <= budget else max(3, min(args.steps, int(budget / est) - 2))
    e2e_warm = args.warmup if e2e_steps == args.steps else min(args.warmup, 2)
    r2 = None
    for _ in range(e2e_warm):
        r2 = None
        r2 = e2e_call()
    barrier()
    t0 = time.perf_counter()
    d2h, e2e_times = 0, []
    for _ in range(e2e_steps):
        r2 = None
        ts = time.perf_counter()
        r2 = e2e_call()
        e2e_times.append(time.perf_counter() - ts)
        d2h = r2.profile["d2h_bytes"]
        if rank == 0:
            log("[e2e] " + json.dumps({k: round(r2.profile[k], 4) for k in
                                       ("t_total", "t_upload", "t_lanczos", "t_ritvec", "t_imtql2", "t_device")}) +
                (" " + json.dumps({k: round(v, 4) for k, v in e2e_parts.items()}) if world > 1 else ""))
    barrier()
    t_e2e = reduce_over_ranks(time.perf_counter() - t0)
    e2e_value = n_opb * e2e_steps / t_e2e
    same = (r2.d == res.d and np.array_equal(r2.s, res.s) and r2.diagnostics == res.diagnostics)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    e2e_sorted = sorted(e2e_times)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h), "time_to_svd_s": t_e2e / e2e_steps, "steps": e2e_steps, "warmup": e2e_warm,
                    "step_s_median": statistics.median(e2e_times), "step_s_p95": e2e_sorted[min(len(e2e_sorted) - 1, int(0.95 * len(e2e_sorted)))],
                    "host_buffers": "pinned in place (cudaHostRegister)" if pinned else "pageable"},
            "time_to_svd_s": {"resident": t_dev / args.steps, "resident_wall": wall_resident / args.steps,
                              "e2e": t_e2e / e2e_steps, "first_upload": upload_first_s},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity": parity, "exchange": exchange,
            "layout": {"n_side_relabelled_by_popularity": relabelled, "gathers_served_from_shared_memory_prefix": hot_fraction},
            "solve": {"lanczos_steps": res.diagnostics.lanczos_steps, "n_opb": int(n_opb), "n_opb_run": int(res.profile["n_opb"]),
                      "n_spmv_run": int(res.profile["n_spmv"]),
                      "ritz_values_stabilized": res.diagnostics.ritz_values_stabilized, "d": res.d,
                      "accepted": res.diagnostics.singular_values, "sigma_max": float(res.s[0]) if res.d else None,
                      "sigma_min": float(res.s[-1]) if res.d else None,
                      "purge_fired": int(res.profile["n_purge_fired"]), "purge_vectors": int(res.profile["n_purge_vectors"]),
                      "host_syncs": int(res.profile["n_host_syncs"]),
                      "phases_s": {k: res.profile[k] for k in ("t_lanczos", "t_sync_wait", "t_ritvec", "t_imtql2", "t_download")},
                      "e2e_phases_s": {k: r2.profile[k] for k in ("t_upload", "t_lanczos", "t_ritvec", "t_download", "t_total")},
                      "e2e_equals_resident": bool(same)}}

    # ---------------- extra: the configuration earlier rounds reported (cfg 2), resident leg only, so the series stays comparable
    if world == 1 and not args.no_extra and W.name != "cfg2":
        try:
            line["extra"] = {"cfg2": extra_workload("cfg2", local, las2, peak)}
        except Exception as e:  # never let the side measurement take the headline line down
            line["extra"] = {"cfg2": {"error": repr(e)}}

    # ---------------- cpu_baseline: the oracle on this box's host cores, bounded sample (rank 0, N == 1 only)
    if world == 1 and not args.no_cpu_baseline:
        counts = reference_counts(fx, res)
        k = max(1, min(10, int(15.0 / max(1e-3, 4.5e-9 * W.nnz))))
        cs = cpu_sample(W, counts, k)
        line["cpu_baseline"] = {"value": cs["opb_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": sample_text(cs, counts), "host_cpus": host_cores(),
                                "sample_detail": {"kind": "extrapolated", "parts_s": cs["parts_s"], "measured": cs["measured"],
                                                  "full_run_on_build_box_s": (fx or {}).get("cpu_seconds", {}).get("total")}}
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", help="cfg1..cfg5 of BASELINE.json (default cfg4 = configs[3], the 1e9-nnz "
                                                       "configuration the north-star targets are quoted on; it fits one B200)")
    ap.add_argument("--scale", type=float, default=1.0, help="scale every dimension (tests only; 1.0 = the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg2 side measurement (N = 1 only)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: fewer than 3 warm-up steps requested; the timing rules ask for >= 3")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
