"""GPU: every compiled variant of k_primal / k_dual, the creation-time tuning that picks one, and the
watchdog of the halo waits.  (Sorted after test_gpu_parity.py on purpose: with `-x` the long-validated
default path is checked first.)"""
import numpy as np
import pytest

from conftest import CASE_PARAMS, case_args

pytestmark = pytest.mark.gpu


def gold_y(g):
    return np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])


@pytest.mark.parametrize("kernel_variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 2 | (4 << 8), 6 | (3 << 8), 8 | (9 << 8)])
@pytest.mark.parametrize("compressed", [False, True])
def test_every_kernel_variant_gives_the_same_bits(kernel_variant, compressed):
    from pysparselp_b200 import _cabi
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    flags = (_cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS) if compressed else 0
    for name in ("potts50", "sc105", "l1svm", "random_small"):
        args, g = case_args(name)
        x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=flags,
                                             kernel_variant=kernel_variant, return_solver=True,
                                             **CASE_PARAMS.get(name, {}))
        try:
            info = solver.info()
            assert info["primal_variant"] == kernel_variant & 0xFF and not info["autotuned"]
            assert np.array_equal(x, g["x_100"]), name
            assert np.array_equal(solver.get_y(), gold_y(g)), name
        finally:
            solver.close()


def test_variants_on_a_mid_size_lp_with_ragged_rows():
    """Potts 256^2 (rows of 3 entries, columns of 2 to 8) and a random LP (rows of 8 entries, equalities, columns
    of ragged length): all variants against the oracle."""
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    cases = [generators.lp_args(generators.potts_lp(256)),
             generators.lp_args(generators.random_sparse_lp(3000, 5000, n_eq=400, seed=7)[0])]
    for args in cases:
        with np.errstate(invalid="ignore"):
            xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=30, nb_iter_plot=1000)
        for v in range(1, 10):
            x, _ = chambolle_pock_ppd(*args, nb_max_iter=30, nb_iter_plot=1000, kernel_variant=v)
            assert np.array_equal(x, xo), v


def test_autotune_keeps_the_initial_state_and_reports_its_timings(monkeypatch):
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    monkeypatch.setenv("CPPPD_AUTOTUNE_MIN_NNZ", "0")
    args, _ = case_args("random_small")
    x0 = np.random.default_rng(3).standard_normal(args[0].size)
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    x, _, solver = chambolle_pock_ppd(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000, return_solver=True)
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo)
    assert info["autotuned"] == 1 and 1 <= info["primal_variant"] <= 9 and 1 <= info["dual_variant"] <= 9
    assert all(v > 0 for v in info["variant_ms"]["k_primal"] + info["variant_ms"]["k_dual"])


def test_autotune_is_on_by_default_for_large_operands():
    """4 M entries and more: the variants are timed on the real operands (Potts 1024^2: 12.6 M entries)."""
    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import make_solver

    lp = generators.potts_lp(1024)
    solver = make_solver(*generators.lp_args(lp))
    try:
        info = solver.info()
        assert info["autotuned"] == 1
        ref = make_solver(*generators.lp_args(lp), kernel_variant=1)
        try:
            solver.iterate(20)
            ref.iterate(20)
            assert np.array_equal(solver.get_x(), ref.get_x()) and np.array_equal(solver.get_y(), ref.get_y())
        finally:
            ref.close()
    finally:
        solver.close()
