"""GPU: the persistent kernels for small LPs (one CTA: k_tiny_iterate; one thread-block cluster: k_cluster_iterate).
(Sorted last on purpose: with `-x` everything on the default path — golden cases, kernel variants, autotune, long rows,
the full-size workloads — is checked first.)"""
import numpy as np
import pytest

from conftest import CASE_PARAMS, case_args

pytestmark = pytest.mark.gpu


def gold_y(g):
    return np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])


# (kb2: 146 distinct values but a CTA of 64 threads — the dictionary is staged in a loop)
@pytest.mark.parametrize("name", ["sc105", "random_small", "random_small_alpha", "kb2", "afiro"])
def test_tiny_persistent_kernel_gives_the_same_bits(name):
    """CPPPD_FLAG_TINY_PERSISTENT: all iterations between two stats blocks in one launch of one CTA."""
    from pysparselp_b200 import _cabi
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd
    from test_gpu_parity import assert_curves_close

    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    for flags in (_cabi.FLAG_TINY_PERSISTENT,
                  _cabi.FLAG_TINY_PERSISTENT | _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS | _cabi.FLAG_REORDER):
        trace = []
        x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=flags, return_solver=True,
                                             callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)),
                                             **kw)
        try:
            assert solver.info()["tiny_persistent"] == 1 and solver.niter == 100
            y = solver.get_y()
            if "alpha" not in kw:
                assert np.array_equal(x, g["x_100"]) and np.array_equal(y, gold_y(g))
            else:
                assert np.allclose(x, g["x_100"], rtol=1e-9, atol=0)
            assert_curves_close(np.array(trace), g["trace_10"])
        finally:
            solver.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["potts50", "sc105", "random_small", "random_small_alpha", "kb2", "afiro", "sc50a", "sc50b"])
def test_cluster_persistent_kernel_gives_the_same_bits(name, mode, monkeypatch):
    """k_cluster_iterate (cpppd_cluster.cuh): all iterations between two stats blocks in one launch of one thread-block
    cluster, operands and vectors in (distributed) shared memory: the default for every LP that fits 16 SMs."""
    from pysparselp_b200 import _cabi
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd
    from test_gpu_parity import assert_curves_close

    # 0: relaxed cluster barrier arrive (after membar.cta); 1: release / acquire
    monkeypatch.setenv("CPPPD_CLUSTER_MODE", str(mode))
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    for flags in (0, _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS | _cabi.FLAG_REORDER):
        trace = []
        x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=flags, return_solver=True,
                                             callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)),
                                             **kw)
        try:
            info = solver.info()
            assert info["tiny_persistent"] == 2 and solver.niter == 100
            y = solver.get_y()
            if "alpha" not in kw:
                assert np.array_equal(x, g["x_100"]) and np.array_equal(y, gold_y(g))
            else:
                assert np.allclose(x, g["x_100"], rtol=1e-9, atol=0)
            assert_curves_close(np.array(trace), g["trace_10"])
        finally:
            solver.close()


def test_cluster_kernel_reproduces_the_potts50_regression_curve():
    """reference tests/test_pott_segmentation.py through SparseLP.solve: the 55 golden curve points, with the
    iterations between two callbacks in one launch of the cluster kernel, and against the CUDA-graph path."""
    import json
    import os
    import time

    from conftest import GOLDEN
    from pysparselp_b200 import _cabi
    from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)
    curves, timings = {}, {}
    for flags in (0, _cabi.FLAG_NO_TINY_PERSISTENT):
        lp, gt, gti, _ = build_linear_program(50, 0.5, 500)
        t0 = time.perf_counter()
        lp.solve(method="chambolle_pock_ppd", get_timing=True, nb_iter=27500, max_time=150, ground_truth=gt,
                 ground_truth_indices=gti, nb_iter_plot=500, flags=flags)
        timings[flags] = time.perf_counter() - t0
        curves[flags] = list(lp.distance_to_ground_truth)
        assert len(curves[flags]) == 55
        np.testing.assert_almost_equal(curves[flags], ref["potts50"][:55])
    assert curves[0] == curves[_cabi.FLAG_NO_TINY_PERSISTENT]
    print("Potts 50x50 regression, 27500 iterations: cluster kernel %.3f s, CUDA graphs %.3f s" % (
        timings[0], timings[_cabi.FLAG_NO_TINY_PERSISTENT]))


def test_tiny_persistent_kernel_reproduces_the_sc105_regression_curve():
    """reference tests/test_netlib.py through SparseLP.solve with the persistent kernel: 41 500 iterations in 83
    launches of k_tiny_iterate (plus the stats blocks)."""
    import json
    import os
    import time

    from conftest import GOLDEN
    from pysparselp_b200 import _cabi
    from pysparselp_b200.netlib import get_problem
    from pysparselp_b200.SparseLP import SparseLP

    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)
    d = get_problem("SC105")
    gt = d["solution"]
    timings = {}
    for flags in (0, _cabi.FLAG_TINY_PERSISTENT):
        lp = SparseLP()
        lp.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                               upper_bounds=np.minimum(d["upper_bounds"], np.max(gt) * 2), costs=d["cost_vector"])
        lp.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
        lp.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
        lp.convert_to_one_sided_inequality_system()
        t0 = time.perf_counter()
        lp.solve(method="chambolle_pock_ppd", get_timing=True, nb_iter=41500, max_time=100, ground_truth=gt,
                 ground_truth_indices=np.arange(len(gt)), nb_iter_plot=500, flags=flags)
        timings[flags] = time.perf_counter() - t0
        assert len(lp.distance_to_ground_truth) == 83
        np.testing.assert_almost_equal(lp.distance_to_ground_truth, ref["SC105"][:83])
    print("SC105, 41500 iterations: CUDA graphs %.3f s, persistent CTA %.3f s" % (timings[0], timings[_cabi.FLAG_TINY_PERSISTENT]))
