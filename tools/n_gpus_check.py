"""One Python process, N GPUs (no torchrun): chambolle_pock_ppd(..., n_gpus=N) and SparseLP.solve(..., n_gpus=N) against
the single-process C port / the reference's Potts regression curve.  python tools/n_gpus_check.py N"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle.c_port import COracle
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

if __name__ == "__main__":
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    for name, lp in (("potts 512", generators.potts_lp(512)), ("random 300k", generators.random_sparse_lp_chunked(300000, 600000)[0])):
        args = generators.lp_args(lp)
        co = COracle(*args)
        co.iterate(40)
        seen = []
        t0 = time.perf_counter()
        x, best = chambolle_pock_ppd(*args, nb_max_iter=40, nb_iter_plot=10, n_gpus=n_gpus,
                                     callback_func=lambda k, xx, e1, e2, el, a, b: seen.append((k, float(np.sum(xx)))))
        dt = time.perf_counter() - t0
        ok = np.array_equal(x, co.x)
        print("%s on %d GPUs from one process: x %s the C port's bits, %d callbacks, %.1f s" % (
            name, n_gpus, "==" if ok else "!=", len(seen), dt), flush=True)
        assert ok and [k for k, _ in seen] == [0, 10, 20, 30]
    # the reference's own regression test (tests/test_pott_segmentation.py) through SparseLP.solve on N GPUs
    from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

    with open(os.path.join(ROOT, "tests", "golden", "reference_curves.json")) as f:
        ref = json.load(f)
    lp, gt, gti, _ = build_linear_program(50, 0.5, 500)
    t0 = time.perf_counter()
    lp.solve(method="chambolle_pock_ppd", get_timing=True, nb_iter=27500, max_time=150, ground_truth=gt,
             ground_truth_indices=gti, nb_iter_plot=500, n_gpus=n_gpus)
    curve = lp.distance_to_ground_truth
    print("Potts 50x50 regression (reference tests/test_pott_segmentation.py) through SparseLP.solve(n_gpus=%d): %d curve "
          "points in %.1f s" % (n_gpus, len(curve), time.perf_counter() - t0), flush=True)
    assert len(curve) == 55
    np.testing.assert_almost_equal(curve, ref["potts50"][: len(curve)])
    print("N_GPUS_CHECK_OK")
