"""Scratch: the long-row kernel shapes / segment orders / segment lengths on one L1-SVM LP (one GPU)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import CpPpdSolver, stack_operator, one_sided_rows

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100000)
ap.add_argument("--features", type=int, default=1000)
a = ap.parse_args()
lp, _ = generators.l1svm_lp(a.size, a.features)
a_in, b_in = one_sided_rows(lp.a_ineq, lp.b_lower, lp.b_upper)
A, b, m_eq = stack_operator(lp.a_eq, lp.b_eq, a_in, b_in, lp.c.size)
ref = None
for seg in (16384, 8192, 4096):
    for order in (1, 0):
        for shape in (0, 1, 2):
            if order == 0 and shape != 0:
                continue
            os.environ.update(CPPPD_LONG_SEG=str(seg), CPPPD_LONG_ORDER=str(order), CPPPD_LONG_SHAPE=str(shape))
            s = CpPpdSolver(lp.c, A, m_eq, b, lp.lb, lp.ub)
            s.iterate(10); s.sync()
            ms = float(np.median([s.time_iterations(20) / 20 for _ in range(3)]))
            kp, kd = [v / 8 for v in s.time_kernels(8)]
            x = s.get_x()
            if seg == 16384:
                ref = x if ref is None else ref
                same = bool(np.array_equal(ref, x))
            else:
                same = float(np.max(np.abs(ref - x)) / max(np.max(np.abs(ref)), 1e-300))
            print(json.dumps(dict(size=a.size, seg=seg, order=order, shape=shape, primal_ms=kp, dual_ms=kd, ms_per_iter=ms,
                                  it_per_s=1e3 / ms, same_x=same)), flush=True)
            del s
