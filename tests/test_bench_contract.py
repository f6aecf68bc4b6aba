"""bench.py's reference arm runs on the CPU (the oracle port, rank 0 only): check the JSON contract the driver reads.
The GPU arm's line carries the same keys plus roofline / gpu_launches / clocks (checked on the GPU box by running it)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra, env=None):
    e = dict(os.environ, **(env or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--scale", "0.005",
                        "--steps", "1", "--warmup", "0", *extra], capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip().splitlines()


def test_reference_arm_line():
    out = run_bench()
    assert len(out) == 1  # exactly one JSON line on stdout
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["unit"] == "opb/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("svd_dim_seed opb matvecs/s") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("cfg2") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "opb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = run_bench("--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out == [] or out == [""]


def test_reference_arm_extrapolated_sample():
    """Workloads with a recorded full-size oracle run (tests/golden/oracle_<cfg>.json) use the SURVEY 8(d) sample: a few real
    svd_opb, one purge sweep, the BLAS-1 chain, contraction terms and imtql2 at the final js, multiplied by the counts of
    the oracle's full solve; the line says so in a structured `sample_detail`."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    out = p.stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["config"]["name"] == "cfg1" and d["value"] > 0
    sd = d["cpu_baseline"]["sample_detail"]
    assert sd["kind"] == "extrapolated" and sd["counts"]["lanczos_steps"] == 782 and sd["opb_timed_total"] >= 10
    assert sd["fixture"].endswith("oracle_cfg1.json") and sd["full_run_on_build_box_s"] > 0
    # the extrapolation reproduces the measured full run of the same box to within a factor of two
    assert 0.5 < d["time_to_svd_s"]["extrapolated"] / sd["full_run_on_build_box_s"] < 2.0
    assert d["e2e"] == {"value": d["value"], "unit": "opb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
