"""Command line around the drop-in call, in the spirit of SVDLIBC's `svd` tool (the reference crate has none):

    python -m svdlibrs_b200 matrix.mtx [-d DIMS] [-i ITERATIONS] [-k KAPPA] [-e LO HI] [-s SEED] [-o PREFIX]

reads a Matrix Market / SVDLIBC sparse-text / dense-text file, runs svdLAS2 on the GPU (there is no CPU fallback) and,
with -o, writes PREFIX-Ut, PREFIX-S, PREFIX-Vt as dense text like the tool's `-w dt`.  Prints the Diagnostics."""
import argparse
import json
import sys
import time

from . import svdLAS2
from . import io as las2_io


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m svdlibrs_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("matrix")
    ap.add_argument("-d", "--dimensions", type=int, default=0, help="number of singular triplets wanted (0 = all)")
    ap.add_argument("-i", "--iterations", type=int, default=0, help="upper bound on Lanczos steps (0 = min(rows, cols))")
    ap.add_argument("-k", "--kappa", type=float, default=1e-6, help="relative accuracy of the Ritz values accepted")
    ap.add_argument("-e", "--end-interval", type=float, nargs=2, default=(-1e-30, 1e-30), metavar=("LO", "HI"))
    ap.add_argument("-s", "--seed", type=int, default=0, help="random seed of the start vector (0 = from entropy)")
    ap.add_argument("-o", "--output", default="", help="prefix of the Ut / S / Vt dense-text files to write")
    a = ap.parse_args(argv)
    A = las2_io.load(a.matrix)
    t0 = time.perf_counter()
    r = svdLAS2(A, a.dimensions, a.iterations, tuple(a.end_interval), a.kappa, a.seed)
    dt = time.perf_counter() - t0
    g = r.diagnostics
    print(json.dumps({"matrix": a.matrix, "shape": list(A.shape), "time_to_svd_s": round(dt, 6), "d": r.d,
                      "singular_values_head": [float(v) for v in r.s[:10]], "diagnostics": g.__dict__}))
    if a.output:
        las2_io.write_svdlibc_dense_text(a.output + "-Ut", r.ut)
        las2_io.write_svdlibc_dense_text(a.output + "-S", r.s)
        las2_io.write_svdlibc_dense_text(a.output + "-Vt", r.vt)
    return 0


if __name__ == "__main__":
    sys.exit(main())
