"""CPU: the WHOLE CUDA library (setup, transpose, SELL build, preconditioners, hot kernels, stats block,
CUDA-graph replay, renumbering, compression) compiled for the host by tests/emul/make_emul.py and driven
through the C ABI with the product's own schedule — the same assertions as tests/test_gpu_parity.py,
without a GPU.  This checks the logic of every kernel and of the host code; what only hardware can show
(memory model, scheduling, performance) stays with the `-m gpu` tests."""
import numpy as np
import pytest

from conftest import CASE_PARAMS, GOLDEN_CASES, case_args
from emul.cabi_driver import emulated_chambolle_pock_ppd, make_emulated_solver
from pysparselp_b200 import _cabi

# (the small goldens would otherwise all run in the persistent CTA, which has its own tests below: keep the
#  thread-per-row kernels + CUDA graphs under test here, and the library default as one more set)
GRAPHS = _cabi.FLAG_NO_TINY_PERSISTENT
FLAG_SETS = {
    "plain": GRAPHS,
    "renumbered": GRAPHS | _cabi.FLAG_REORDER,
    "compressed": GRAPHS | _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS,
    "all": GRAPHS | _cabi.FLAG_REORDER | _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS,
    "default": 0,
}


def rel_inf(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def curves_close(got, want):
    assert got.shape == want.shape and np.array_equal(got[:, 0], want[:, 0])
    for col in range(1, want.shape[1]):
        g, w = got[:, col], want[:, col]
        assert np.array_equal(np.isnan(g), np.isnan(w))
        inf = np.isinf(w)
        assert np.array_equal(g[inf], w[inf])
        fin = np.isfinite(w)
        floor = 1e-6 * max(np.max(np.abs(w[fin])) if fin.any() else 0.0, 1e-30)
        assert np.all(np.abs(g[fin] - w[fin]) <= 1e-6 * np.abs(w[fin]) + (floor if col in (1, 2) else 0.0))


@pytest.mark.parametrize("variant", list(FLAG_SETS))
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_library_vs_golden(name, variant):
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=FLAG_SETS[variant], partition_granule=32,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
    try:
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        s_gold = np.concatenate([g[k] for k in ("diag_sigma_eq", "diag_sigma_ineq") if k in g])
        assert solver.niter == 100
        assert np.array_equal(T == 1.0, g["diag_t"] == 1.0)
        if "alpha" not in kw:
            assert np.array_equal(T, g["diag_t"]) and np.array_equal(sigma, s_gold)
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold)
        else:
            assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        curves_close(np.array(trace), g["trace_10"])
        assert (best is None) == (g["best_100"].size == 0)
    finally:
        solver.close()


@pytest.mark.parametrize("name", ["potts50", "sc105"])
def test_emulated_force_integer_bookkeeping(name):
    args, g = case_args(name)
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=300, nb_iter_plot=20, force_integer=True,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)))
    solver.close()
    assert np.array_equal(x, g["x_300_fi"])
    curves_close(np.array(trace), g["trace_20_fi"])
    assert best is not None and np.array_equal(best, g["best_300_fi"])


def test_emulated_partition_matches_python_restatement():
    import scipy.sparse as sp

    from oracle import partition_oracle as po
    from pysparselp_b200 import generators

    for lp, m_eq in ((generators.potts_lp(20), 0), (generators.random_sparse_lp(300, 400, n_eq=70, seed=4)[0], 70)):
        solver = make_emulated_solver(*generators.lp_args(lp), flags=_cabi.FLAG_REORDER, partition_granule=32)
        cols, ghost_c = solver.layout(columns=True)
        rows, ghost_r = solver.layout(columns=False)
        info = solver.info()
        solver.close()
        blocks = [a for a in (lp.a_eq, lp.a_ineq) if a is not None]
        a = sp.vstack(blocks).tocsr() if len(blocks) > 1 else blocks[0]
        part = po.partition(a.indptr, a.indices, a.shape[1], m_eq, 1, granule=32)
        assert ghost_c.size == 0 and ghost_r.size == 0
        assert np.array_equal(cols, part["col_order"]) and np.array_equal(rows, part["row_order"])
        assert info["nnz"] == a.nnz and info["a_padded_entries"] >= a.nnz


def test_emulated_automatic_renumbering_on_padding():
    """A pattern whose row lengths vary wildly is renumbered automatically (padding > 15 %), results unchanged."""
    import scipy.sparse as sp

    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    rng = np.random.default_rng(11)
    n, m = 200, 260
    rows = []
    for i in range(m):
        k = 1 if i % 7 else 25
        cols = rng.choice(n, size=k, replace=False)
        rows.append((cols, np.round(rng.standard_normal(k), 1) + 0.05))
    indptr = np.concatenate(([0], np.cumsum([r[0].size for r in rows])))
    a = sp.csr_matrix((np.concatenate([r[1] for r in rows]), np.concatenate([r[0] for r in rows]), indptr), shape=(m, n))
    c = rng.standard_normal(n)
    lb, ub = -np.ones(n), np.ones(n)
    args = (c, sp.csr_matrix((0, n)), np.empty(0), a, None, rng.random(m), lb, ub)
    xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=40, nb_iter_plot=10**6)
    # one locality bucket (granule >= n): inside it rows / columns are grouped by length (sigma sorting)
    x, _, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=40, nb_iter_plot=10**6, partition_granule=4096)
    cols, _ = solver.layout(columns=True)
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo)
    assert not np.array_equal(cols, np.arange(n)), "heavy padding should have triggered the renumbering"
    padded_sorted = info["a_padded_entries"] + info["at_padded_entries"]
    x2, _, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=40, nb_iter_plot=10**6, flags=_cabi.FLAG_NO_REORDER)
    cols2, _ = solver.layout(columns=True)
    padded_identity = sum(solver.info()[k] for k in ("a_padded_entries", "at_padded_entries"))
    solver.close()
    assert np.array_equal(x2, xo) and np.array_equal(cols2, np.arange(n))
    assert padded_identity > 1.15 * 2 * a.nnz and padded_sorted < 0.7 * padded_identity


def test_emulated_abi_errors():
    import ctypes as C

    import scipy.sparse as sp

    from emul.cabi_driver import emulated_library

    lib = emulated_library()
    a = sp.csr_matrix(np.array([[1.0, -1.0], [0.5, 2.0]]))
    with pytest.raises(_cabi.CpppdError, match="column index"):
        bad = a.copy()
        bad.indices = np.array([0, 5, 0, 1], dtype=np.int32)
        make_emulated_solver(np.ones(2), None, None, bad, None, np.zeros(2), np.zeros(2), np.ones(2))
    s = make_emulated_solver(np.ones(2), None, None, a, None, np.zeros(2), np.zeros(2), np.ones(2))
    with pytest.raises(_cabi.CpppdError, match="primal step"):
        s.dual_step()
    with pytest.raises(_cabi.CpppdError):
        s.stats_step()
    s.primal_step(keep_d=True)
    with pytest.raises(_cabi.CpppdError, match="dual_step"):
        s.iterate(3)
    s.close()
    assert lib.cpppd_destroy(None) == 0


@pytest.mark.parametrize("case", ["potts50", "sc105"])
def test_solve_curves_on_device_equal_host_curves(case, monkeypatch):
    """SparseLP.solve(): the per-callback curves evaluated inside the stats block (x stays on the device)
    equal the reference-style host evaluation (x downloaded at every callback), and the reference goldens."""
    import json
    import os

    import pysparselp_b200.ChambollePockPPD as front
    from conftest import GOLDEN
    from emul.patch_plugin import _Adapter

    monkeypatch.setattr(front, "CpPpdSolver", _Adapter)
    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)

    def build():
        if case == "potts50":
            from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

            lp, gt, gti, _ = build_linear_program(50, 0.5, 500)
            return lp, gt, gti, ref["potts50"]
        from pysparselp_b200.netlib import get_problem
        from pysparselp_b200.SparseLP import SparseLP

        d = get_problem("SC105")
        gt = d["solution"]
        lp = SparseLP()
        lp.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                               upper_bounds=np.minimum(d["upper_bounds"], np.max(gt) * 2), costs=d["cost_vector"])
        lp.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
        lp.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
        lp.convert_to_one_sided_inequality_system()
        return lp, gt, np.arange(len(gt)), ref["SC105"]

    curves = {}
    for device_curves in (True, False):
        lp, gt, gti, golden = build()
        x, _ = lp.solve(method="chambolle_pock_ppd", nb_iter=1001, nb_iter_plot=500, ground_truth=gt,
                        ground_truth_indices=gti, device_curves=device_curves)
        curves[device_curves] = (x, {k: np.array(getattr(lp, k), dtype=float) for k in (
            "distance_to_ground_truth", "distanceToGroundTruthAfterRounding", "pobj_curve", "dobj_curve",
            "max_violated_constraint", "max_violated_equality", "max_violated_inequality", "itrn_curve")})
    (xd, cd), (xh, ch) = curves[True], curves[False]
    assert np.array_equal(xd, xh)
    for k in ch:
        if k.startswith("distance"):
            assert np.allclose(cd[k], ch[k], rtol=1e-13, atol=1e-15), k
        else:
            assert np.array_equal(cd[k], ch[k], equal_nan=True), k
    np.testing.assert_almost_equal(cd["distance_to_ground_truth"], golden[: len(cd["distance_to_ground_truth"])])


@pytest.mark.parametrize("kernel_variant", [1, 2, 3, 4, 5, 6, 7, 3 | (5 << 8), 7 | (6 << 8)])
@pytest.mark.parametrize("flags", [0, _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS])
def test_emulated_every_kernel_variant_gives_the_same_bits(kernel_variant, flags):
    """cpppd_problem.kernel_variant forces one of the compiled variants of k_primal / k_dual; the arithmetic of
    a row / column sum is the same sequential chain in all of them."""
    for name in ("sc105", "random_small"):  # (tests/test_kernels_on_cpu.py runs every variant on every golden case)
        args, g = case_args(name)
        trace = []
        x, best, solver = emulated_chambolle_pock_ppd(
            *args, nb_max_iter=100, nb_iter_plot=10, flags=flags, kernel_variant=kernel_variant,
            callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)))
        info = solver.info()
        y = solver.get_y()
        solver.close()
        assert info["primal_variant"] == kernel_variant & 0xFF
        assert info["dual_variant"] == ((kernel_variant >> 8) & 0xFF or kernel_variant & 0xFF)
        assert not info["autotuned"]
        assert np.array_equal(x, g["x_100"])
        assert np.array_equal(y, np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g]))
        curves_close(np.array(trace), g["trace_10"])


def test_emulated_bad_kernel_variant_is_refused():
    args, _ = case_args("sc105")
    with pytest.raises(_cabi.CpppdError, match="kernel_variant"):
        make_emulated_solver(*args, kernel_variant=_cabi.KERNEL_VARIANTS + 1)


def test_emulated_autotune_leaves_the_initial_state_untouched(monkeypatch):
    """Timing the variants at creation runs the kernels on the real operands: x0 / xbar / y must be put back."""
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    monkeypatch.setenv("CPPPD_AUTOTUNE_MIN_NNZ", "0")
    args, g = case_args("random_small")
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(args[0].size)
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    x, _, solver = emulated_chambolle_pock_ppd(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo)
    assert info["autotuned"] == 1 and 1 <= info["primal_variant"] <= 7 and 1 <= info["dual_variant"] <= 7
    assert all(v > 0 for v in info["variant_ms"]["k_primal"]) and all(v > 0 for v in info["variant_ms"]["k_dual"])
    # not timed when asked not to
    x, _, solver = emulated_chambolle_pock_ppd(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000, flags=_cabi.FLAG_NO_AUTOTUNE)
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo) and info["autotuned"] == 0 and info["primal_variant"] == 1


def test_emulated_autotune_choice_is_remembered_per_operand_shape(monkeypatch):
    """A second solve on operands of the same shape reuses the measured choice (same timings reported, to the
    bit); CPPPD_AUTOTUNE_CACHE=0 measures again; another shape is measured on its own."""
    monkeypatch.setenv("CPPPD_AUTOTUNE_MIN_NNZ", "0")
    args, g = case_args("sc105")

    def tuned_info(a, **kw):
        x, _, solver = emulated_chambolle_pock_ppd(*a, nb_max_iter=100, nb_iter_plot=10, **kw)
        info = solver.info()
        solver.close()
        assert np.array_equal(x, g["x_100"]) and info["autotuned"] == 1
        return info

    first, second = tuned_info(args), tuned_info(args)
    assert second["variant_ms"] == first["variant_ms"]
    assert (second["primal_variant"], second["dual_variant"]) == (first["primal_variant"], first["dual_variant"])
    # SC105 pads by more than 15 % and is renumbered automatically; the caller's numbering gives other operands
    other_shape = tuned_info(args, flags=_cabi.FLAG_NO_REORDER)
    assert other_shape["variant_ms"] != first["variant_ms"]
    monkeypatch.setenv("CPPPD_AUTOTUNE_CACHE", "0")
    again = tuned_info(args)
    assert again["variant_ms"] != first["variant_ms"]  # wall-clock timings of a new measurement


# thresholds that make a handful of rows / columns long (the emulator runs one CTA of fibers per segment: slow)
LONG_THRESHOLD = {"l1svm": 64, "sc105": 3, "random_small": 13}


@pytest.mark.parametrize("variant", ["plain", "all"])
@pytest.mark.parametrize("name", list(LONG_THRESHOLD))
def test_emulated_long_rows_agree_to_rounding(name, variant):
    """Rows / columns above cpppd_problem.long_row_threshold leave the thread-per-row kernels and are summed by a
    CTA per segment (fixed tree): iterates within BASELINE.json's 1e-9, curves within 1e-6, masks exact."""
    args, g = case_args(name)
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=FLAG_SETS[variant], partition_granule=32,
        long_row_threshold=LONG_THRESHOLD[name], callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)))
    try:
        info = solver.info()
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        s_gold = np.concatenate([g[k] for k in ("diag_sigma_eq", "diag_sigma_ineq") if k in g])
        assert info["long_rows"] + info["long_cols"] > 0
        assert info["long_entries"] > LONG_THRESHOLD[name] * (info["long_rows"] + info["long_cols"])
        assert np.array_equal(T == 1.0, g["diag_t"] == 1.0) and np.array_equal(sigma == 1.0, s_gold == 1.0)
        assert rel_inf(T, g["diag_t"]) < 1e-14 and rel_inf(sigma, s_gold) < 1e-14
        assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        curves_close(np.array(trace), g["trace_10"])
        assert (best is None) == (g["best_100"].size == 0)
    finally:
        solver.close()


def test_emulated_long_rows_with_segments_and_force_integer():
    """Rows longer than one 4096-entry segment (several CTAs per row), a dense 'budget' row and a dense column,
    against the oracle; the default threshold leaves the golden cases on the exact path."""
    import scipy.sparse as sp

    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    rng = np.random.default_rng(21)
    n, m = 9000, 60
    a = sp.random(m, n, density=0.002, random_state=5, format="lil")
    a[7, :] = np.round(rng.standard_normal(n), 2) + 0.005      # 9000 entries: three segments
    a[:, 11] = (np.round(rng.standard_normal(m), 2) + 0.005)[:, None]
    a = a.tocsr()
    a_eq = sp.csr_matrix(np.ones((1, n)))                      # dense equality row
    xf = rng.random(n)
    c = np.round(rng.standard_normal(n), 2)
    lb, ub = np.zeros(n), np.ones(n)
    args = (c, a_eq, a_eq @ xf, a, None, a @ xf + 0.1, lb, ub)
    tr_o, tr = [], []
    with np.errstate(invalid="ignore"):
        xo, bo = chambolle_pock_ppd_oracle(*args, nb_max_iter=60, nb_iter_plot=20, force_integer=True,
                                           callback_func=lambda k, xx, e1, e2, el, p, q: tr_o.append((k, e1, e2, p, q)))
    x, best, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=60, nb_iter_plot=20, force_integer=True,
                                                  callback_func=lambda k, xx, e1, e2, el, p, q: tr.append((k, e1, e2, p, q)))
    info = solver.info()
    solver.close()
    assert info["long_rows"] == 2 and info["long_cols"] == 0 and info["long_entries"] == 2 * n  # default threshold 2048
    assert rel_inf(x, xo) <= 1e-9
    curves_close(np.array(tr), np.array(tr_o))
    assert (best is None) == (bo is None)
    # never split: bit-identical again
    x2, _, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=60, nb_iter_plot=20, force_integer=True,
                                                long_row_threshold=-1)
    assert solver.info()["long_rows"] == 0
    solver.close()
    assert np.array_equal(x2, xo)
    args_g, g = case_args("l1svm")
    x3, _, solver = emulated_chambolle_pock_ppd(*args_g, nb_max_iter=100, nb_iter_plot=10)
    assert solver.info()["long_cols"] == 0
    solver.close()
    assert np.array_equal(x3, g["x_100"])


# (kb2: 146 distinct values but a CTA of 64 threads — the dictionary is staged in a loop)
@pytest.mark.parametrize("name", ["sc105", "random_small", "random_small_alpha", "kb2", "afiro"])
@pytest.mark.parametrize("flags", [_cabi.FLAG_TINY_PERSISTENT,
                                   _cabi.FLAG_TINY_PERSISTENT | _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS | _cabi.FLAG_REORDER])
def test_emulated_tiny_persistent_kernel(name, flags):
    """CPPPD_FLAG_TINY_PERSISTENT: the iterations between two stats blocks run in one launch of one CTA."""
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=flags,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
    try:
        assert solver.info()["tiny_persistent"] == 1 and solver.niter == 100
        y = solver.get_y()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold)
        else:
            assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        curves_close(np.array(trace), g["trace_10"])
    finally:
        solver.close()
    # too large for one CTA: the cluster kernel takes it; with CPPPD_FLAG_NO_TINY_PERSISTENT the graph path
    args, g = case_args("potts50")
    for flag, want in ((_cabi.FLAG_TINY_PERSISTENT, 2), (_cabi.FLAG_NO_TINY_PERSISTENT, 0)):
        x, _, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=20, nb_iter_plot=10, flags=flag)
        assert solver.info()["tiny_persistent"] == want
        solver.close()


@pytest.mark.parametrize("name", ["potts50", "sc105", "random_small", "random_small_alpha", "kb2", "afiro", "sc50a", "sc50b"])
@pytest.mark.parametrize("flags", [0, _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS | _cabi.FLAG_REORDER])
def test_emulated_cluster_kernel(name, flags):
    """k_cluster_iterate (csrc/cpppd_cluster.cuh), the default for small LPs: its staging (index words translated into
    [owner CTA][offset], slices padded to whole chunks with value-0 entries) and its per-row code compiled for the host
    and driven phase by phase over one buffer per CTA — what the hardware cluster does between two cluster barriers.
    Same bits as the goldens; the one-CTA kernel and the graph path give the same x, y."""
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=flags,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
    try:
        assert solver.info()["tiny_persistent"] == 2 and solver.niter == 100
        y = solver.get_y()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold)
        else:
            assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        curves_close(np.array(trace), g["trace_10"])
    finally:
        solver.close()
    x0, _, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10,
                                                flags=flags | _cabi.FLAG_NO_TINY_PERSISTENT, **kw)
    try:
        assert solver.info()["tiny_persistent"] == 0
        assert np.array_equal(x0, x) and np.array_equal(solver.get_y(), y)
    finally:
        solver.close()


@pytest.mark.parametrize("window", ["small", "single"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_banded_operands_give_the_same_bits(name, window, monkeypatch):
    """CPPPD_FLAG_BANDED (cpppd_banded.cuh): window-major operands, one launch per window of the gathered vector,
    partial row sums carried in memory between windows.  Rows with ascending column indices keep their summation
    order, so the iterates are the golden bits for every window size (7 elements: dozens of windows per pass;
    100000: a single window).  Operands whose rows are not window-ordered must stay with the SELL kernels."""
    if name == "l1svm":  # weight columns hold ~1 350 entries: a window may hold at most 255 of them (one count byte)
        window = 199 if window == "small" else 251
    else:
        window = 100000 if window == "single" else (997 if name == "potts50" else 7)
    monkeypatch.setenv("CPPPD_BAND_WINDOW", str(window))
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    trace = []
    x, best, solver = emulated_chambolle_pock_ppd(
        *args, nb_max_iter=100, nb_iter_plot=10, flags=_cabi.FLAG_BANDED,
        callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
    try:
        info = solver.info()
        y = solver.get_y()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        # A^T is window-ordered by construction (columns sorted by source row): always banded
        assert info["band_in_use"][1] == 1 and info["band_windows"][1] >= 1
        if window < 100000:
            assert info["band_windows"][1] > 2
        else:  # one window per kind of row: equality and inequality duals are never mixed in a window
            assert info["band_windows"][0] <= 1 and info["band_windows"][1] <= 2
        if name in ("sc105", "random_small", "afiro", "kb2"):  # rows with ascending column indices: A as well
            assert info["band_in_use"][0] == 1
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold)
        else:
            assert rel_inf(x, g["x_100"]) <= 1e-9 and rel_inf(y, y_gold) <= 1e-9
        curves_close(np.array(trace), g["trace_10"])
    finally:
        solver.close()


@pytest.mark.parametrize("name,window", [("random_small", 11), ("l1svm", 251)])
@pytest.mark.parametrize("shape", [1, 2, 3, 4, 5, 6, 7])
def test_emulated_banded_kernel_shapes_give_the_same_bits(shape, name, window, monkeypatch):
    """The window kernels are compiled in eight shapes — entries through registers (flat entries per lane and trip /
    CTAs per SM) or staged in shared memory by bulk copies (gathers in flight per lane) — and cpppd_create times
    them on large operands.  Every shape must produce the golden bits (shape 0 is what the other banded tests run).
    The L1-SVM case has tiles of ~32 000 entries per window: dozens of staged pieces of 384 entries per tile."""
    monkeypatch.setenv("CPPPD_BAND_WINDOW", str(window))
    monkeypatch.setenv("CPPPD_BAND_SHAPE", str(shape))
    args, g = case_args(name)
    x, best, solver = emulated_chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=_cabi.FLAG_BANDED)
    try:
        info = solver.info()
        assert info["band_in_use"][1] == 1 and info["band_shape"][1] == shape
        if name == "random_small":
            assert info["band_in_use"] == [1, 1]
        assert np.array_equal(x, g["x_100"])
        assert np.array_equal(solver.get_y(), np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g]))
    finally:
        solver.close()


def test_solve_with_fixed_variables_keeps_x_on_the_device(monkeypatch):
    """``SparseLP.solve`` eliminates fixed variables before the solve and records, at every callback, curves of the
    FULL LP at the mapped-back point (reference SparseLP.py:632-674, :1064-1093).  With the per-row offsets of
    ``cpppd_set_row_offsets`` and the constant terms of the eliminated entries these are evaluated on the device: x is
    fetched once, at the end — and the curves equal the host-side evaluation (``device_curves=False``)."""
    import pysparselp_b200.ChambollePockPPD as front
    from emul.patch_plugin import _Adapter
    from pysparselp_b200.SparseLP import SparseLP
    from test_fuzz_on_cpu import random_lp

    fetches = []

    class Counting(_Adapter):
        def get_x(self):
            fetches.append(1)
            return super().get_x()

    monkeypatch.setattr(front, "CpPpdSolver", Counting)
    args, _, _, rng = random_lp(412)
    c, a_eq, b_eq, a_in, b_lo, b_up, lb, ub = args
    n = c.size
    lb = np.where(np.isfinite(lb), lb, -3.0)
    ub = np.where(np.isfinite(ub), ub, 3.0)
    fix = rng.random(n) < 0.35
    ub = np.where(fix, lb, ub)
    gt_idx = np.sort(rng.choice(n, size=n // 2, replace=False))
    gt = np.round(rng.standard_normal(gt_idx.size), 1)
    curves = {}
    for device_curves in (True, False):
        lp = SparseLP()
        lp.add_variables_array(n, lower_bounds=lb, upper_bounds=ub, costs=c)
        if a_eq is not None and a_eq.shape[0]:
            lp.add_equality_constraints_sparse(a_eq, b_eq)
        lp.add_inequality_constraints_sparse(a_in, b_lo if b_lo is not None else np.full(a_in.shape[0], -np.inf), b_up)
        fetches.clear()
        x, _ = lp.solve(method="chambolle_pock_ppd", nb_iter=60, nb_iter_plot=10, ground_truth=gt, ground_truth_indices=gt_idx,
                        device_curves=device_curves)
        assert len(fetches) == (1 if device_curves else 7)
        curves[device_curves] = (x, {k: np.array(getattr(lp, k), dtype=float) for k in (
            "distance_to_ground_truth", "distanceToGroundTruthAfterRounding", "max_violated_constraint",
            "max_violated_equality", "max_violated_inequality", "pobj_curve", "dobj_curve")})
    assert fix.sum() > 0 and np.array_equal(curves[True][0], curves[False][0])
    for k, want in curves[False][1].items():
        got = curves[True][1][k]
        assert got.shape == want.shape == (6,)
        fin = np.isfinite(want)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
        assert np.allclose(got[fin], want[fin], rtol=1e-9, atol=1e-12), k
