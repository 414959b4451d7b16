"""(Named to sort after test_gpu_parity.py: with `-x` the long-validated default path is checked first.)
GPU: LPs with rows / columns too long for one thread (cpppd_long_rows.cuh): summed by a CTA per segment in a
fixed tree, so they agree with the reference to rounding — BASELINE.json's 1e-9 on the iterates, 1e-6 on the
curves, exact "replaced by 1" masks — while LPs without such rows stay bit-identical."""
import numpy as np
import pytest

from conftest import case_args

pytestmark = pytest.mark.gpu

LONG_THRESHOLD = {"l1svm": 64, "sc105": 3, "random_small": 13}


def rel_inf(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("flags", [0, 11])  # 11 = renumbering + value dictionary + constant vectors
@pytest.mark.parametrize("name", list(LONG_THRESHOLD))
def test_long_rows_agree_with_the_goldens_to_rounding(name, flags):
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd
    from test_gpu_parity import assert_curves_close

    args, g = case_args(name)
    trace = []
    x, best, solver = chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=flags, long_row_threshold=LONG_THRESHOLD[name], return_solver=True,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)))
    try:
        info = solver.info()
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        s_gold = np.concatenate([g[k] for k in ("diag_sigma_eq", "diag_sigma_ineq") if k in g])
        assert info["long_rows"] + info["long_cols"] > 0
        assert np.array_equal(T == 1.0, g["diag_t"] == 1.0) and np.array_equal(sigma == 1.0, s_gold == 1.0)
        assert rel_inf(T, g["diag_t"]) < 1e-14 and rel_inf(sigma, s_gold) < 1e-14
        assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        assert_curves_close(np.array(trace), g["trace_10"])
    finally:
        solver.close()


def test_l1svm_shaped_lp_with_the_default_threshold():
    """L1-SVM, 6000 samples x 8 features (BASELINE configs[2] in small): the 27 weight columns hold 8000 entries
    each — two segments per column with the default threshold — everything else stays on the thread-per-row path."""
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd
    from test_gpu_parity import assert_curves_close

    lp, _ = generators.l1svm_lp(6000, 8)
    args = generators.lp_args(lp)
    tr_o, tr = [], []
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=40, nb_iter_plot=10,
                                          callback_func=lambda k, xx, e1, e2, el, a, b: tr_o.append((k, e1, e2, a, b)))
    x, _, solver = chambolle_pock_ppd(*args, nb_max_iter=40, nb_iter_plot=10, return_solver=True,
                                      callback_func=lambda k, xx, e1, e2, el, a, b: tr.append((k, e1, e2, a, b)))
    info = solver.info()
    solver.close()
    assert info["long_cols"] == 27 and info["long_rows"] == 0 and info["long_entries"] > 27 * 7000
    assert rel_inf(x, xo) <= 1e-9
    assert_curves_close(np.array(tr), np.array(tr_o))
    # never split: the same LP bit for bit
    x2, _ = chambolle_pock_ppd(*args, nb_max_iter=40, nb_iter_plot=10, long_row_threshold=-1)
    assert np.array_equal(x2, xo)


def test_dense_budget_row_and_dense_column():
    """A 9000-entry inequality row (three segments), a dense equality row, a dense column, force_integer."""
    import scipy.sparse as sp

    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd
    from test_gpu_parity import assert_curves_close

    rng = np.random.default_rng(21)
    n, m = 9000, 60
    a = sp.random(m, n, density=0.002, random_state=5, format="lil")
    a[7, :] = np.round(rng.standard_normal(n), 2) + 0.005
    a[:, 11] = (np.round(rng.standard_normal(m), 2) + 0.005)[:, None]
    a = a.tocsr()
    a_eq = sp.csr_matrix(np.ones((1, n)))
    xf = rng.random(n)
    c = np.round(rng.standard_normal(n), 2)
    args = (c, a_eq, a_eq @ xf, a, None, a @ xf + 0.1, np.zeros(n), np.ones(n))
    tr_o, tr = [], []
    with np.errstate(invalid="ignore"):
        xo, bo = chambolle_pock_ppd_oracle(*args, nb_max_iter=60, nb_iter_plot=20, force_integer=True,
                                           callback_func=lambda k, xx, e1, e2, el, p, q: tr_o.append((k, e1, e2, p, q)))
    x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=60, nb_iter_plot=20, force_integer=True, return_solver=True,
                                         callback_func=lambda k, xx, e1, e2, el, p, q: tr.append((k, e1, e2, p, q)))
    info = solver.info()
    solver.close()
    assert info["long_rows"] == 2 and info["long_entries"] == 2 * n
    assert rel_inf(x, xo) <= 1e-9
    assert_curves_close(np.array(tr), np.array(tr_o))
    assert (best is None) == (bo is None)
