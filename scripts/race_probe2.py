import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVDLAS2_BLOCKED"] = "2"
import svdlibrs_b200 as las2
from svdlibrs_b200 import synthetic
A, _ = synthetic.make("cfg2", scale=1.0)
rng = np.random.default_rng(0)
with las2.Operator(A) as op:
    for tma, transposed in ((1, True), (1, False), (1, True), (1, False), (1, True)):
        op.set('spmv_tma', tma)
        x = rng.standard_normal(A.shape[0] if transposed else A.shape[1])
        op.set("spmv_variant", 2); ref = op.opa(x, transposed)
        for seq in ([4, 4, 4, 4, 4, 4, 4, 4], ):
            out = []; LASTBAD = globals().get('LASTBAD', [])
            for v in seq:
                op.set("spmv_variant", v)
                y = op.opa(x, transposed)
                bad = np.flatnonzero(~(np.abs(y - ref) <= 1e-9 * (1 + np.abs(ref))))
                LASTBAD = bad[:6].tolist() if (bad.size and transposed) else LASTBAD; globals()['LASTBAD'] = LASTBAD
                out.append((v, int(bad.size), int(np.isnan(y).sum()), bad[:6].tolist()))
            print("tma", tma, "transposed", transposed, "seq", out, flush=True)
    op.set("spmv_variant", -1)
    import scipy.sparse as sp
    AT = A.T.tocsr()
    rl = np.diff(AT.indptr)
    print("BT row lengths around a bad row:", [(int(r), int(rl[r]), int(AT.indptr[r]) ) for r in (LASTBAD if LASTBAD else [])])
