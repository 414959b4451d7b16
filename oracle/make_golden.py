"""TEST INFRASTRUCTURE — mint golden vectors from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For each case it
  1. builds the LP with the reference's own modeling code and with this repo's builders and
     asserts that the solver inputs are array-equal (so tests can rebuild them offline),
  2. runs the reference ``chambolle_pock_ppd`` (``pysparselp/ChambollePockPPD.py:36-346``) and
     records, per stats block, ``(niter, energy1, energy2, max_violated_equality,
     max_violated_inequality)`` and the callback x; the final ``(x, best_integer_solution)``;
     and — read from the reference's stack frame inside its callback, because the function
     does not return them — ``y_eq, y_ineq, x3, diag_t, diag_sigma_*`` after 100 iterations,
  3. runs ``oracle/cpppd_oracle.py`` on the same inputs and asserts it reproduces all of the
     above bit for bit (this is what pins the oracle),
  4. writes ``tests/golden/<case>.npz`` (+ a sha256 of the solver inputs).

It also copies the CP-PPD entries of the reference's own golden files into
``tests/golden/reference_curves.json``.
"""
import contextlib
import copy
import hashlib
import io
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.cpppd_oracle import chambolle_pock_ppd_oracle  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def lp_digest(args):
    """sha256 over the solver inputs (c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub)."""
    h = hashlib.sha256()
    for a in args:
        if a is None:
            h.update(b"none")
        elif sp.issparse(a):
            a = sp.csr_matrix(a)
            h.update(np.asarray(a.shape, dtype=np.int64).tobytes())
            h.update(np.ascontiguousarray(a.indptr, dtype=np.int64).tobytes())
            h.update(np.ascontiguousarray(a.indices, dtype=np.int64).tobytes())
            h.update(np.ascontiguousarray(a.data, dtype=np.float64).tobytes())
        else:
            h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def solver_args_from_lp(lp):
    """What ``SparseLP.solve`` hands to the solver (reference ``SparseLP.py:1244-1287``)."""
    red = copy.deepcopy(lp)
    red.remove_fixed_variables()
    return (red.costsvector, red.a_equalities, red.b_equalities, red.a_inequalities, red.b_lower,
            red.b_upper, red.lower_bounds, red.upper_bounds)


def assert_same_args(a, b):
    assert len(a) == len(b)
    for u, v in zip(a, b):
        if u is None or v is None:
            assert u is None and v is None
        elif sp.issparse(u):
            u, v = sp.csr_matrix(u), sp.csr_matrix(v)
            assert u.shape == v.shape, (u.shape, v.shape)
            assert np.array_equal(u.indptr, v.indptr)
            assert np.array_equal(u.indices, v.indices)
            assert np.array_equal(u.data, v.data)
        else:
            assert np.array_equal(np.asarray(u), np.asarray(v))


def run_traced(fn, args, nb_max_iter, nb_iter_plot, force_integer=False, grab_frame=False, **kw):
    trace, xs, frame_vars = [], [], {}

    def cb(niter, x, e1, e2, elapsed, mv_eq, mv_ineq):
        trace.append((niter, e1, e2, mv_eq, mv_ineq))
        xs.append(x.copy())
        if grab_frame:
            loc = sys._getframe(1).f_locals
            for name in ("y_eq", "y_ineq", "x3", "diag_t", "diag_sigma_eq", "diag_sigma_ineq", "d"):
                if name in loc and loc[name] is not None:
                    frame_vars[name] = np.array(loc[name], dtype=np.float64, copy=True)

    with contextlib.redirect_stdout(io.StringIO()):
        out = fn(*args, nb_max_iter=nb_max_iter, nb_iter_plot=nb_iter_plot, callback_func=cb,
                 force_integer=force_integer, **kw)
    return out, np.array(trace, dtype=np.float64), xs, frame_vars


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def mint(name, args, extra=None, alpha=1, theta=1):
    ref = ref_loader.reference_chambolle_pock_ppd()
    kw = dict(alpha=alpha, theta=theta)
    out = {}
    # run A: 100 iterations, stats every 10
    (x_ref, best_ref), tr_ref, xs_ref, _ = run_traced(ref, args, 100, 10, **kw)
    (x_or, best_or), tr_or, xs_or, _ = run_traced(chambolle_pock_ppd_oracle, args, 100, 10, **kw)
    assert same(x_ref, x_or), name
    assert same(best_ref, best_or), name
    assert same(tr_ref, tr_or), (name, np.nanmax(np.abs(tr_ref - tr_or)))
    assert all(same(a, b) for a, b in zip(xs_ref, xs_or))
    out.update(x_100=x_ref, trace_10=tr_ref, x_at_50=xs_ref[5])
    out["best_100"] = best_ref if best_ref is not None else np.zeros(0)
    # run B: state after exactly 100 iterations, lifted from the reference's frame at niter 100
    _, _, _, fv = run_traced(ref, args, 101, 100, grab_frame=True, **kw)
    state = {}
    chambolle_pock_ppd_oracle(*args, nb_max_iter=100, nb_iter_plot=10, state_out=state, **kw)
    for key_ref, key_or in (("y_eq", "y_eq"), ("y_ineq", "y_ineq"), ("diag_t", "diag_t"),
                            ("diag_sigma_eq", "sig_eq"), ("diag_sigma_ineq", "sig_ineq")):
        if key_ref in fv:
            assert same(fv[key_ref], state[key_or]), (name, key_ref)
            out[key_ref] = fv[key_ref]
    # run C: force_integer bookkeeping, 300 iterations, stats every 20
    (x_ref, best_ref), tr_ref, _, _ = run_traced(ref, args, 300, 20, force_integer=True, **kw)
    (x_or, best_or), tr_or, _, _ = run_traced(chambolle_pock_ppd_oracle, args, 300, 20, force_integer=True, **kw)
    assert same(x_ref, x_or) and same(best_ref, best_or) and same(tr_ref, tr_or), name
    out.update(x_300_fi=x_ref, trace_20_fi=tr_ref)
    out["best_300_fi"] = best_ref if best_ref is not None else np.zeros(0)
    out["digest"] = np.frombuffer(lp_digest(args).encode(), dtype=np.uint8)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("minted %-12s n=%d  callbacks=%d  best@100=%s best@300fi=%s" % (
        name, args[0].size, len(tr_ref), out["best_100"].size > 0, out["best_300_fi"].size > 0))


def pack_lp(args):
    c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub = args
    d = dict(lp_c=c, lp_lb=lb, lp_ub=ub)
    for tag, a in (("eq", a_eq), ("ineq", a_ineq)):
        if a is not None:
            a = sp.csr_matrix(a)
            d["lp_a%s_indptr" % tag] = a.indptr
            d["lp_a%s_indices" % tag] = a.indices
            d["lp_a%s_data" % tag] = a.data
            d["lp_a%s_shape" % tag] = np.asarray(a.shape)
    if beq is not None:
        d["lp_beq"] = beq
    if b_lower is not None:
        d["lp_b_lower"] = b_lower
    if b_upper is not None:
        d["lp_b_upper"] = b_upper
    return d


def main(only=()):
    """Mint every golden, or only the named cases (``python -m oracle.make_golden kb2 sc50a``)."""
    os.makedirs(GOLDEN, exist_ok=True)
    ref_loader.load_reference()
    import importlib

    def want(name):
        return not only or name in only

    from pysparselp_b200 import generators
    from pysparselp_b200.examples import example_l1_svm as my_svm
    from pysparselp_b200.examples import example_pott_segmentation as my_potts
    from pysparselp_b200.netlib import get_problem as my_get_problem
    from pysparselp_b200.SparseLP import SparseLP as MySparseLP

    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- C1: Potts 50x50 (reference example_pott_segmentation.py:54-92)
    ref_potts = importlib.import_module("pysparselp.examples.example_pott_segmentation")
    with quiet:
        lp_ref, gt_ref, gti_ref, _ = ref_potts.build_linear_program(50, 0.5, 500)
    lp_my, gt_my, gti_my, _ = my_potts.build_linear_program(50, 0.5, 500)
    args = solver_args_from_lp(lp_ref)
    assert_same_args(args, solver_args_from_lp(lp_my))
    assert np.array_equal(gt_ref, gt_my) and np.array_equal(gti_ref, gti_my)
    g = generators.potts_lp(50)
    assert_same_args(args, (g.c, sp.csr_matrix((0, g.c.size)), np.empty(0), g.a_ineq,
                            np.full(g.b_upper.size, -np.inf), g.b_upper, g.lb, g.ub))
    if want("potts50"):
        mint("potts50", args, extra=dict(ground_truth=gt_ref))

    # ---- C2: netlib SC105 (reference tests/test_netlib.py:19-48)
    ref_netlib = importlib.import_module("pysparselp.netlib")
    ref_slp = ref_loader.reference_sparse_lp()

    def netlib_lp(get_problem, cls, problem):
        d = get_problem(problem)
        gt = d["solution"]
        lp = cls()
        lp.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                               upper_bounds=np.minimum(d["upper_bounds"], np.max(gt) * 2),
                               costs=d["cost_vector"])
        lp.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
        lp.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
        lp.convert_to_one_sided_inequality_system()
        return lp, gt

    # SC105 is the case the reference tests; the other problems it vendors (AFIRO, KB2 — the one with a BOUNDS
    # section —, SC50A, SC50B) go through the same preparation: MPS parser -> modeling layer -> solver
    for problem in ("SC105", "AFIRO", "KB2", "SC50A", "SC50B"):
        if not want(problem.lower()):
            continue
        with quiet:
            lp_ref, gt_ref = netlib_lp(ref_netlib.get_problem, ref_slp.SparseLP, problem)
        # (only SC105 and AFIRO are vendored in this package: the others are parsed from the reference's data folder)
        ref_data = os.path.join(ref_loader.REFERENCE_ROOT, "pysparselp", "data")
        lp_my, gt_my = netlib_lp(lambda name: my_get_problem(name, data_dir=ref_data), MySparseLP, problem)
        args = solver_args_from_lp(lp_ref)
        assert_same_args(args, solver_args_from_lp(lp_my))
        assert np.array_equal(gt_ref, gt_my)
        mint(problem.lower(), args, extra=dict(ground_truth=gt_ref, **pack_lp(args)))

    # ---- L1-SVM 1000 x 2 (reference example_l1_svm.py:91-113)
    ref_svm = importlib.import_module("pysparselp.examples.example_l1_svm")
    x, classes = my_svm.make_data()
    svm_ref = ref_svm.L1SVM()
    svm_ref.set_data(x, classes)
    svm_my = my_svm.L1SVM()
    svm_my.set_data(x, classes)
    args = solver_args_from_lp(svm_ref)
    assert_same_args(args, solver_args_from_lp(svm_my))
    g, _ = generators.l1svm_lp(1000, 2)
    assert_same_args(args, (g.c, sp.csr_matrix((0, g.c.size)), np.empty(0), g.a_ineq, g.b_lower, g.b_upper,
                            g.lb, g.ub))
    if want("l1svm"):
        mint("l1svm", args)

    # ---- small random LP with equalities, two-sided rows, infinite bounds and an x0-free start
    rng = np.random.default_rng(7)
    n, me, mi = 120, 25, 160
    a_eq = sp.random(me, n, density=0.06, random_state=11, format="csr")
    a_eq.data = np.round(a_eq.data * 8 - 4, 2)
    a_in = sp.random(mi, n, density=0.05, random_state=12, format="csr")
    a_in.data = np.round(a_in.data * 8 - 4, 2)
    xf = np.round(rng.standard_normal(n), 2)
    beq = a_eq @ xf
    mid = a_in @ xf
    b_lo = mid - np.abs(np.round(rng.standard_normal(mi), 2))
    b_up = mid + np.abs(np.round(rng.standard_normal(mi), 2))
    b_lo[rng.random(mi) < 0.4] = -np.inf
    b_up[rng.random(mi) < 0.3] = np.inf
    t = np.round(rng.standard_normal(n), 2)
    lb = xf + np.minimum(0, t) - 0.5
    ub = xf + np.maximum(0, t) + 0.5
    lb[rng.random(n) < 0.1] = -np.inf
    ub[rng.random(n) < 0.1] = np.inf
    c = np.round(rng.standard_normal(n), 2)
    args = (c, a_eq, beq, a_in, b_lo, b_up, lb, ub)
    if want("random_small"):
        mint("random_small", args, extra=pack_lp(args))
    if want("random_small_alpha"):
        mint("random_small_alpha", args, extra=pack_lp(args), alpha=1.5, theta=0.7)

    # ---- the reference's own goldens for the path
    if not want("reference_curves"):
        return
    curves = {}
    for key, fname in (("SC105", "netlib_curves_SC105.json"), ("potts50", "test_pott_segmentation_curves.json"),
                       ("l1svm", "test_l1_svm_results.json")):
        with open(os.path.join(ref_loader.REFERENCE_ROOT, "tests", fname)) as f:
            curves[key] = json.load(f)["chambolle_pock_ppd"]
    with open(os.path.join(GOLDEN, "reference_curves.json"), "w") as f:
        json.dump(curves, f)
    print("copied reference goldens:", {k: (len(v) if isinstance(v, list) else v) for k, v in curves.items()})


if __name__ == "__main__":
    main(only=set(sys.argv[1:]))
