"""Scratch timing of the inner loop on one GPU (not the bench contract; see bench.py)."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import CpPpdSolver, stack_operator, one_sided_rows

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--kind", default="potts")
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--variant", type=int, default=0, help="kernel_variant (0: autotune); primal | dual << 8")
a = ap.parse_args()
t0 = time.time()
if a.kind == "potts":
    lp = generators.potts_lp(a.size)
elif a.kind == "random":
    lp, _ = generators.random_sparse_lp_chunked(a.size, 2 * a.size, nnz_per_row=8)
elif a.kind == "l1svm":  # --size samples x 1000 features (BASELINE configs[2] family)
    lp, _ = generators.l1svm_lp(a.size, 1000)
t1 = time.time()
a_in, b_in = one_sided_rows(lp.a_ineq, lp.b_lower, lp.b_upper)
A, b, m_eq = stack_operator(lp.a_eq, lp.b_eq, a_in, b_in, lp.c.size)
s = CpPpdSolver(lp.c, A, m_eq, b, lp.lb, lp.ub, flags=a.flags, kernel_variant=a.variant)
t2 = time.time()
info = s.info()
s.iterate(20); s.sync()
times = [s.time_iterations(a.iters) / a.iters for _ in range(a.reps)]
ms = float(np.median(times))
kp, kd = [v / 16 for v in s.time_kernels(16)]
out = dict(kind=a.kind, size=a.size, variant=a.variant, chosen=(info["primal_variant"], info["dual_variant"]),
           primal_ms=kp, dual_ms=kd, n=info["n"], m=info["m_eq"] + info["m_ineq"], nnz=info["nnz"],
           build_s=round(t1 - t0, 2), setup_s=round(t2 - t1, 2), ms_per_iter=ms, it_per_s=1e3 / ms,
           algo_GBs=info["bytes_per_iteration_algorithmic"] / ms / 1e6,
           actual_GBs=info["bytes_per_iteration_actual"] / ms / 1e6, times=times,
           pad_A=info["a_padded_entries"] / max(info["nnz"], 1), pad_AT=info["at_padded_entries"] / max(info["nnz"], 1),
           device_GB=info["device_bytes"] / 1e9, long_rows=info["long_rows"], long_cols=info["long_cols"],
           long_entries=info["long_entries"], band_windows=info["band_windows"], band_in_use=info["band_in_use"],
           band_ms=info["band_ms"], band_shape=info["band_shape"], band_shape_ms=info["band_shape_ms"], band_spg=info["band_sectors_per_gather"], band_window_mb=info["band_window_bytes"] / 2**20,
           variant_ms=info["variant_ms"], band_env=os.environ.get("CPPPD_BAND_WINDOW_MB"))
print(json.dumps(out))
