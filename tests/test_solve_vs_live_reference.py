"""CPU, build container only: ``SparseLP.solve(method='chambolle_pock_ppd')`` of this package (CUDA library on the
CPU emulator) against the UNMODIFIED reference's ``SparseLP.solve`` run live, on random LPs built through the
modeling layer of each side with the same calls — variable blocks, sparse equality / inequality rows, optional
one-sided conversion, fixed variables (``remove_fixed_variables``, the ``- shift`` mapping of ``SparseLP.py:1259,
:1288``), ground-truth curves.  Skipped where /root/reference does not exist.
"""
import contextlib
import io

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_loader
from test_fuzz_on_cpu import random_lp

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")

CURVES = ("distance_to_ground_truth", "distanceToGroundTruthAfterRounding", "pobj_curve", "dobj_curve",
          "max_violated_constraint", "max_violated_equality", "max_violated_inequality", "itrn_curve")


def build(cls, args, one_sided, split):
    c, a_eq, b_eq, a_in, b_lo, b_up, lb, ub = args
    n = c.size
    lp = cls()
    if split and n > 1:  # two variable blocks
        k = n // 2
        lp.add_variables_array(k, lower_bounds=lb[:k], upper_bounds=ub[:k], costs=c[:k])
        lp.add_variables_array(n - k, lower_bounds=lb[k:], upper_bounds=ub[k:], costs=c[k:])
    else:
        lp.add_variables_array(n, lower_bounds=lb, upper_bounds=ub, costs=c)
    if a_eq is not None and a_eq.shape[0]:
        lp.add_equality_constraints_sparse(a_eq, b_eq)
    lp.add_inequality_constraints_sparse(a_in, b_lo if b_lo is not None else np.full(a_in.shape[0], -np.inf), b_up)
    if one_sided:
        lp.convert_to_one_sided_inequality_system()
    return lp


@pytest.mark.parametrize("seed", range(300, 330))
def test_solve_equals_the_live_reference(seed, monkeypatch):
    import pysparselp_b200.ChambollePockPPD as front
    from emul.patch_plugin import _Adapter
    from pysparselp_b200.SparseLP import SparseLP as Mine

    monkeypatch.setattr(front, "CpPpdSolver", _Adapter)
    args, _, _, rng = random_lp(seed)
    if args[3] is None:
        pytest.skip("the reference fails without an inequality block (ChambollePockPPD.py:283)")
    c, lb, ub = args[0], args[6], args[7]
    n = c.size
    finite = np.isfinite(lb) & np.isfinite(ub)
    fix = finite & (rng.random(n) < 0.3)  # fixed variables: remove_fixed_variables has work to do
    ub = ub.copy()
    ub[fix] = lb[fix]
    args = args[:7] + (ub,)
    one_sided, split = bool(rng.random() < 0.5), bool(rng.random() < 0.5)
    gt_idx = np.sort(rng.choice(n, size=max(1, n // 2), replace=False))
    gt = np.round(rng.standard_normal(gt_idx.size), 1)
    kw = dict(method="chambolle_pock_ppd", nb_iter=int(rng.integers(1, 80)), nb_iter_plot=int(rng.choice([1, 5, 20])),
              ground_truth=gt, ground_truth_indices=gt_idx)
    theirs = build(ref_loader.reference_sparse_lp().SparseLP, args, one_sided, split)
    mine = build(Mine, args, one_sided, split)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        x_r, _ = theirs.solve(**kw)
    with np.errstate(all="ignore"):
        x_m, _ = mine.solve(**kw)
        x_h, _ = build(Mine, args, one_sided, split).solve(device_curves=False, **kw)
    assert np.array_equal(x_m, x_r, equal_nan=True) and np.array_equal(x_h, x_r, equal_nan=True)
    for k in CURVES:
        r, m = np.array(getattr(theirs, k), dtype=float), np.array(getattr(mine, k), dtype=float)
        assert r.shape == m.shape, k
        assert np.array_equal(np.isnan(r), np.isnan(m)) and np.array_equal(r[np.isinf(r)], m[np.isinf(r)]), k
        fin = np.isfinite(r)
        scale = max(np.max(np.abs(r[fin])) if fin.any() else 0.0, 1e-30)
        assert np.all(np.abs(r[fin] - m[fin]) <= 1e-6 * np.abs(r[fin]) + 1e-9 * scale), (k, r, m)


def _same_matrix(a, b):
    if a is None or b is None:
        return (a is None or a.shape[0] == 0) and (b is None or b.shape[0] == 0)
    a, b = sp.csr_matrix(a), sp.csr_matrix(b)
    a.sort_indices()
    b.sort_indices()
    return a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) \
        and np.array_equal(a.data, b.data)


def _same_vector(a, b):
    if a is None or b is None:
        return (a is None or np.size(a) == 0) and (b is None or np.size(b) == 0)
    return np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


@pytest.mark.parametrize("conversion", ["convert_to_all_equalities", "convert_to_all_inequalities",
                                        "convert_to_all_inequalities_without_bounds",
                                        "convert_to_one_sided_inequality_system"])
@pytest.mark.parametrize("seed", range(400, 412))
def test_modeling_layer_conversions_equal_the_live_reference(seed, conversion):
    from pysparselp_b200.SparseLP import SparseLP as Mine

    args, _, _, rng = random_lp(seed)
    if args[3] is None:
        pytest.skip("needs an inequality block")
    theirs = build(ref_loader.reference_sparse_lp().SparseLP, args, False, bool(seed % 2))
    mine = build(Mine, args, False, bool(seed % 2))
    with contextlib.redirect_stdout(io.StringIO()):
        getattr(theirs, conversion)()
    getattr(mine, conversion)()
    assert mine.nb_variables == theirs.nb_variables
    for name in ("costsvector", "lower_bounds", "upper_bounds", "b_equalities", "b_lower", "b_upper"):
        assert _same_vector(getattr(mine, name), getattr(theirs, name)), name
    assert _same_matrix(mine.a_equalities, theirs.a_equalities), "a_equalities"
    assert _same_matrix(mine.a_inequalities, theirs.a_inequalities), "a_inequalities"
