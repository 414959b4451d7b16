"""CPU: the oracle (oracle/cpppd_oracle.py) against the goldens minted from the unmodified
reference (tests/golden/*.npz, oracle/make_golden.py) and against the reference's own golden
files for the path (tests/golden/reference_curves.json)."""
import json
import os

import numpy as np
import pytest

from conftest import CASE_PARAMS, GOLDEN, GOLDEN_CASES, case_args
from oracle.cpppd_oracle import chambolle_pock_ppd_oracle


def run(args, nb_max_iter, nb_iter_plot, **kw):
    trace, state = [], {}

    def cb(niter, x, e1, e2, elapsed, mv_eq, mv_ineq):
        trace.append((niter, e1, e2, mv_eq, mv_ineq))

    with np.errstate(invalid="ignore"):
        x, best = chambolle_pock_ppd_oracle(*args, nb_max_iter=nb_max_iter, nb_iter_plot=nb_iter_plot,
                                            callback_func=cb, state_out=state, **kw)
    return x, best, np.array(trace), state


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_bit_exact_vs_reference_golden(name):
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    x, best, trace, st = run(args, 100, 10, **kw)
    assert np.array_equal(x, g["x_100"])
    assert np.array_equal(trace, g["trace_10"], equal_nan=True)
    assert (best is None) == (g["best_100"].size == 0)
    for key, val in (("y_eq", st["y_eq"]), ("y_ineq", st["y_ineq"]), ("diag_t", st["diag_t"]),
                     ("diag_sigma_eq", st["sig_eq"]), ("diag_sigma_ineq", st["sig_ineq"])):
        if key in g:
            assert np.array_equal(val, g[key]), key
    x, best, trace, _ = run(args, 300, 20, force_integer=True, **kw)
    assert np.array_equal(x, g["x_300_fi"])
    assert np.array_equal(trace, g["trace_20_fi"], equal_nan=True)
    if g["best_300_fi"].size:
        assert np.array_equal(best, g["best_300_fi"])
    else:
        assert best is None


def _curve_through_solve_semantics(args, gt, gt_idx, nb_iter, nb_iter_plot):
    """distance_to_ground_truth as SparseLP.solve records it (reference SparseLP.py:1074-1077)."""
    curve = []

    def cb(niter, x, *rest):
        curve.append(float(np.mean(np.abs(gt - x[gt_idx]))))

    with np.errstate(invalid="ignore"):
        chambolle_pock_ppd_oracle(*args, nb_max_iter=nb_iter, nb_iter_plot=nb_iter_plot, callback_func=cb)
    return curve


def test_oracle_reproduces_reference_sc105_curve():
    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)["SC105"]
    args, g = case_args("sc105")
    gt = g["ground_truth"]
    curve = _curve_through_solve_semantics(args, gt, np.arange(gt.size), 10001, 500)
    np.testing.assert_almost_equal(curve, ref[: len(curve)])  # the reference's own tolerance (7 decimals)
    assert np.max(np.abs(np.array(curve) - np.array(ref[: len(curve)]))) == 0.0


def test_oracle_reproduces_reference_potts_curve():
    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)["potts50"]
    args, g = case_args("potts50")
    gt = g["ground_truth"]
    idx = np.arange(2500).reshape(50, 50, 1)
    curve = _curve_through_solve_semantics(args, gt, idx, 5001, 500)
    np.testing.assert_almost_equal(curve, ref[: len(curve)])
    assert np.max(np.abs(np.array(curve) - np.array(ref[: len(curve)]))) == 0.0


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_port_bit_exact_vs_reference_golden(name):
    """oracle/cpppd_oracle.c (multi-threaded CPU baseline) against the reference-minted goldens."""
    from oracle.c_port import COracle

    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    o = COracle(*args, **kw)
    o.iterate(100)
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    if "alpha" in kw:  # pow() of libm vs numpy may differ in the last place
        assert np.allclose(o.x, g["x_100"], rtol=1e-12, atol=0) and np.allclose(o.y, y_gold, rtol=1e-12, atol=1e-300)
    else:
        assert np.array_equal(o.x, g["x_100"]) and np.array_equal(o.y, y_gold)
        assert np.array_equal(o.T, g["diag_t"])


@pytest.mark.parametrize("case", ["SC105", "potts50"])
def test_c_port_reproduces_every_point_of_the_reference_curves(case):
    """All 83 (SC105, 41 500 iterations) / 55 (Potts 50x50, 27 500 iterations) points of the reference's own golden
    curves for this path (tests/netlib_curves_SC105.json, tests/test_pott_segmentation_curves.json), through the
    plain-C port: the callback of a stats iteration k sees x after the primal half of iteration k (:242-329)."""
    from oracle.c_port import COracle

    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = np.array(json.load(f)[case])
    args, g = case_args("sc105" if case == "SC105" else "potts50")
    gt = g["ground_truth"]
    idx = np.arange(gt.size) if case == "SC105" else np.arange(2500).reshape(50, 50, 1)
    o = COracle(*args)
    curve = []
    for _ in range(len(ref)):
        o.primal_step()
        curve.append(float(np.mean(np.abs(gt - o.x[idx]))))
        o.dual_step()
        o.iterate(499)
    np.testing.assert_almost_equal(curve, ref)  # the reference's own tolerance (7 decimals)
    assert np.max(np.abs(np.array(curve) - ref)) == 0.0
