"""CPU, build container only: the numpy oracle against the UNMODIFIED reference run live
(/root/reference through oracle/ref_loader.py) on the random LPs of tests/test_fuzz_on_cpu.py.

The committed goldens pin the oracle on nine LPs; this widens the pin to shapes the goldens do not contain (few
columns, empty and all-zero rows, unsorted indices, infinite and fixed bounds, two-sided rows, x0, theta != 1,
alpha != 1, force_integer).  Skipped where the reference tree does not exist (the GPU box): nothing else in the
suite reads /root/reference at run time.
"""
import contextlib
import io

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_loader
from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
from test_fuzz_on_cpu import random_lp

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def traced(fn, args, **kw):
    trace, xs = [], []

    def cb(niter, x, e1, e2, elapsed, mv_eq, mv_ineq):
        trace.append((niter, e1, e2, mv_eq, mv_ineq))
        xs.append(np.array(x, copy=True))

    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        x, best = fn(*args, callback_func=cb, **kw)
    return x, best, np.array(trace, dtype=np.float64), xs


@pytest.mark.parametrize("seed", range(200, 240))
def test_oracle_equals_the_live_reference(seed):
    args, x0, theta, rng = random_lp(seed)
    if args[3] is None:
        pytest.skip("the reference fails without an inequality block (ChambollePockPPD.py:283)")
    alpha = float(rng.choice([1.0, 1.0, 1.0, 1.5, 0.5]))
    kw = dict(x0=x0, theta=theta, alpha=alpha, nb_max_iter=int(rng.integers(1, 60)),
              nb_iter_plot=int(rng.choice([1, 3, 10, 1000])), force_integer=bool(rng.random() < 0.5))
    ref = ref_loader.reference_chambolle_pock_ppd()
    ref_args = args
    if args[1] is None:  # the reference needs a matrix object (solve() hands it a 0-row CSR, :70-72 turns it into None)
        ref_args = (args[0], sp.csr_matrix((0, args[0].size)), np.empty(0)) + args[3:]
    x_r, best_r, trace_r, xs_r = traced(ref, ref_args, **kw)
    x_o, best_o, trace_o, xs_o = traced(chambolle_pock_ppd_oracle, args, **kw)
    assert np.array_equal(x_r, x_o, equal_nan=True)
    assert (best_r is None) == (best_o is None) and (best_r is None or np.array_equal(best_r, best_o))
    assert np.array_equal(trace_r, trace_o, equal_nan=True)
    assert len(xs_r) == len(xs_o) and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(xs_r, xs_o))
