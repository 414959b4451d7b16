"""CPU: host-side logic of the drop-in (modeling layer, MPS parser, conversions, generators)
and the C-ABI library surface (loads, exports every declared symbol, refuses to run without a GPU)."""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, solver_args_from_lp


def test_library_exports_every_declared_symbol():
    from pysparselp_b200 import _cabi, build

    build.build_library()
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "cpppd.h")).read()
    declared = set(re.findall(r"\b(cpppd_[a-z_]+)\s*\(", header))
    declared -= {"cpppd_alloc_fn", "cpppd_free_fn"}
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cpppd_abi_version() == _cabi.ABI_VERSION


def test_struct_layout_matches_header():
    from pysparselp_b200 import _cabi

    # spot checks of the C layout (x86-64 SysV): sizes computed by hand from include/cpppd.h
    assert C.sizeof(_cabi.Stats) == 8 + 13 * 8 + 4 * 4
    # windows, in_use, ms, sectors, window bytes, dense_halo + reserved, shape, shape_ms
    banded = 2 * 4 + 2 * 4 + 2 * 4 + 2 * 4 + 8 + 2 * 4 + 2 * 4 + 2 * 8 * 4
    assert C.sizeof(_cabi.Info) == 9 * 8 + 6 * 4 + 9 * 8 + 2 * 4 + 2 * _cabi.KERNEL_VARIANTS * 4 + 3 * 8 + 2 * 4 + banded
    assert _cabi.Problem.indptr.offset == 40 and _cabi.Problem.alpha.offset == 112
    assert _cabi.Problem.rank.offset == 176 and C.sizeof(_cabi.Problem) == 224  # (+ band_window)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    a = sp.csr_matrix(np.array([[1.0, -1.0]]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chambolle_pock_ppd(np.ones(2), sp.csr_matrix((0, 2)), np.empty(0), a, None, np.zeros(1),
                           np.zeros(2), np.ones(2))


def test_create_reports_missing_device_through_the_abi():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pysparselp_b200 import _cabi

    lib = _cabi.load_library()
    indptr = np.array([0, 2], dtype=np.int32)
    indices = np.array([0, 1], dtype=np.int32)
    vals = np.array([1.0, -1.0])
    v = np.zeros(2)
    p = _cabi.Problem()
    p.abi_version = _cabi.ABI_VERSION
    p.n, p.m_eq, p.m_ineq, p.nnz = 2, 0, 1, 2
    p.indptr, p.indices, p.values = indptr.ctypes.data, indices.ctypes.data, vals.ctypes.data
    p.indptr_bits = p.index_bits = 32
    p.c = p.lb = p.ub = p.b = v.ctypes.data
    h = C.c_void_p()
    rc = lib.cpppd_create(C.byref(p), C.byref(h))
    assert rc == -4 and b"no CPU path" in lib.cpppd_last_error(None)
    p.abi_version = 99
    assert lib.cpppd_create(C.byref(p), C.byref(h)) == -1


def test_one_sided_rows_order_and_values():
    from oracle.cpppd_oracle import one_sided_system
    from pysparselp_b200.ChambollePockPPD import one_sided_rows

    a = sp.csr_matrix(np.arange(1, 13, dtype=float).reshape(4, 3))
    lo = np.array([-np.inf, 1.0, -np.inf, 2.0])
    up = np.array([5.0, np.inf, 6.0, 7.0])
    a2, b2 = one_sided_rows(a, lo, up)
    ao, bo = one_sided_system(a, lo, up)
    assert np.array_equal(a2.toarray(), ao.toarray()) and np.array_equal(b2, bo)
    assert np.array_equal(b2, [5.0, 6.0, 7.0, -1.0, -2.0])
    assert np.array_equal(a2.toarray(), np.vstack((a.toarray()[[0, 2, 3]], -a.toarray()[[1, 3]])))
    # b_lower None: passthrough (reference ChambollePockPPD.py:87-88)
    a3, b3 = one_sided_rows(a, None, up)
    assert a3 is a and b3 is up
    # only lower bounds finite
    a4, b4 = one_sided_rows(a, np.array([1.0, 2, 3, 4]), np.full(4, np.inf))
    assert np.array_equal(a4.toarray(), -a.toarray()) and np.array_equal(b4, [-1, -2, -3, -4])


def test_modeling_layer_basics():
    from pysparselp_b200.SparseLP import SparseLP, crd_matrix

    m = crd_matrix(np.array([[2, 0], [1, 2]]), np.array([[1.0, -1.0], [0.0, 3.0]]))
    assert np.array_equal(m.indices, [2, 0, 2]) and np.array_equal(m.indptr, [0, 2, 3])  # order kept, zero dropped
    with pytest.raises(ValueError):
        crd_matrix(np.array([[1, 1]]), np.array([[1.0, 2.0]]))
    lp = SparseLP()
    ids = lp.add_variables_array((2, 2), 0, 1, costs=np.arange(4.0).reshape(2, 2), name="v")
    aux = lp.add_variables_array(3, None, None)
    assert np.array_equal(lp.get_variables_indices("v"), ids) and aux[0] == 4
    assert np.all(np.isinf(lp.lower_bounds[4:])) and lp.nb_variables == 7
    lp.add_inequality_constraints(np.array([[0, 4]]), np.array([[1.0, -1.0]]), lower_bounds=None, upper_bounds=0)
    lp.add_equality_constraints(np.array([[1, 5]]), np.array([[1.0, 1.0]]), 2)  # scalar bound -> equality block
    lp.add_equality_constraints(np.array([[2, 6]]), np.array([[1.0, 1.0]]), np.array([3.0]))  # array -> two-sided row
    assert lp.nb_equality_constraints() == 1 and lp.nb_inequality_constraints() == 2
    assert np.array_equal(lp.b_lower, [-np.inf, 3.0]) and np.array_equal(lp.b_upper, [0.0, 3.0])
    lp.convert_to_one_sided_inequality_system()
    assert lp.b_lower is None and lp.a_inequalities.shape == (3, 7)
    assert np.array_equal(lp.b_upper, [0.0, 3.0, -3.0])
    x = np.array([0.5, 1.0, 1.0, 0.0, 0.5, 1.0, 2.0])
    assert lp.check_solution(x) and lp.max_constraint_violation(x) == 0
    with pytest.raises(NotImplementedError):
        lp.solve(method="admm")
    with pytest.raises(ValueError):
        lp.solve(method="nope")


def test_remove_fixed_variables():
    from pysparselp_b200.SparseLP import SparseLP

    lp = SparseLP()
    lp.add_variables_array(4, np.array([0.0, 2.0, 0.0, -1.0]), np.array([1.0, 2.0, 3.0, -1.0]), costs=np.arange(4.0))
    lp.add_inequality_constraints_sparse(sp.csr_matrix(np.array([[1.0, 1, 1, 1], [0, 2, 0, 1]])), None, np.array([4.0, 9]))
    lp.add_equality_constraints_sparse(sp.csr_matrix(np.array([[1.0, 0, 0, 3]])), np.array([1.0]))
    m_change, shift = lp.remove_fixed_variables()
    assert lp.nb_variables == 2 and np.array_equal(shift, [0, 2, 0, -1])
    assert np.array_equal(lp.b_upper, [3.0, 6.0]) and np.array_equal(lp.b_equalities, [4.0])
    assert np.array_equal(lp.a_inequalities.toarray(), [[1, 1], [0, 0]])
    assert np.array_equal(m_change.toarray(), [[1, 0], [0, 0], [0, 1], [0, 0]])


MPS_TEXT = """* tiny model
NAME          TESTPROB
ROWS
 N  COST
 L  LIM1
 G  LIM2
 E  MYEQN
COLUMNS
    XONE      COST                 1   LIM1                 1
    XONE      LIM2                 1
    YTWO      COST                 4   LIM1                 1
    YTWO      MYEQN               -1
    ZTHREE    COST                 9   LIM2                 1
    ZTHREE    MYEQN                1
RHS
    RHS1      LIM1                 5   LIM2                10
    RHS1      MYEQN                7
BOUNDS
 UP BND1      XONE                 4
 LO BND1      YTWO                -1
 UP BND1      YTWO                 1
 FR BND1      ZTHREE
ENDATA
"""


def test_mps_parser_small():
    from pysparselp_b200.MPSparser import mps_parser

    d = mps_parser(io.StringIO(MPS_TEXT))
    assert d["problem_name"] == "TESTPROB" and d["costname"] == "COST"
    assert np.array_equal(d["cost_vector"], [1, 4, 9])
    assert np.array_equal(d["lower_bounds"], [0, -1, -np.inf]) and np.array_equal(d["upper_bounds"], [4, 1, np.inf])
    assert np.array_equal(d["a_ineq"].toarray(), [[1, 1, 0], [1, 0, 1]])
    assert np.array_equal(d["b_lower"], [-np.inf, 10]) and np.array_equal(d["b_upper"], [5, np.inf])
    assert np.array_equal(d["a_eq"].toarray(), [[0, -1, 1]]) and np.array_equal(d["b_eq"], [7])
    with pytest.raises(NotImplementedError):
        mps_parser(io.StringIO(MPS_TEXT.replace("BOUNDS", "RANGES")))


def test_netlib_sc105_shapes_and_solution():
    from pysparselp_b200.netlib import get_problem

    d = get_problem("SC105")
    assert d["a_eq"].shape == (45, 103) and d["a_ineq"].shape == (60, 103)
    assert d["a_eq"].nnz == 122 and d["a_ineq"].nnz == 158
    assert np.count_nonzero(d["cost_vector"]) == 1 and d["solution"].shape == (103,)
    x = d["solution"]
    assert np.allclose(d["a_eq"] @ x, d["b_eq"], atol=1e-9) and np.all(d["a_ineq"] @ x <= d["b_upper"] + 1e-9)
    assert abs(d["cost_vector"] @ x - (-5064062500 / 97008861)) < 1e-9
    with pytest.raises(FileNotFoundError):
        get_problem("NOPE")


@pytest.mark.parametrize("problem", ["SC105", "AFIRO", "KB2", "SC50A", "SC50B"])
def test_netlib_problems_reach_the_solver_as_the_reference_builds_them(problem):
    """MPS parser -> modeling layer -> one-sided conversion -> remove_fixed_variables, as reference
    tests/test_netlib.py:19-48 prepares a netlib LP: the solver inputs must hash to the digest recorded when the
    golden was minted from the unmodified reference (oracle/make_golden.py asserted array equality then)."""
    from conftest import load_golden, solver_args_from_lp
    from oracle.make_golden import lp_digest
    from pysparselp_b200.netlib import get_problem
    from pysparselp_b200.SparseLP import SparseLP

    data_dir = None
    if problem not in ("SC105", "AFIRO"):  # not vendored here: parsed from the reference's data folder where it exists
        data_dir = "/root/reference/pysparselp/data"
        if not os.path.isdir(data_dir):
            pytest.skip("reference data folder not present")
    d = get_problem(problem, data_dir=data_dir)
    gt = d["solution"]
    lp = SparseLP()
    lp.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                           upper_bounds=np.minimum(d["upper_bounds"], np.max(gt) * 2), costs=d["cost_vector"])
    lp.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
    lp.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
    lp.convert_to_one_sided_inequality_system()
    assert lp.check_solution(gt)
    g = load_golden(problem.lower())
    assert lp_digest(solver_args_from_lp(lp)) == bytes(g["digest"]).decode()
    assert np.array_equal(gt, g["ground_truth"])


@pytest.mark.parametrize("size", [7, 50])
def test_potts_generator_matches_modeling_layer(size):
    from pysparselp_b200 import generators
    from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

    lp, _, _, _ = build_linear_program(size, 0.5, 500, with_ground_truth=False)
    c, a_eq, beq, a_in, b_lo, b_up, lb, ub = solver_args_from_lp(lp)
    g = generators.potts_lp(size)
    assert np.array_equal(g.c, c) and np.array_equal(g.lb, lb) and np.array_equal(g.ub, ub)
    assert np.array_equal(g.a_ineq.indptr, a_in.indptr) and np.array_equal(g.a_ineq.indices, a_in.indices)
    assert np.array_equal(g.a_ineq.data, a_in.data) and np.array_equal(g.b_upper, b_up)
    assert np.all(b_lo == -np.inf) and a_eq.shape[0] == 0


def test_potts_generator_rectangular_structure():
    from pysparselp_b200 import generators

    g = generators.potts_lp(5, 9)
    P, nh, nv = 45, 5 * 8, 4 * 9
    assert g.c.size == P + nh + nv and g.a_ineq.shape == (2 * (nh + nv), P + nh + nv)
    a = g.a_ineq.toarray()
    assert np.all(np.abs(a).sum(axis=1) == 3) and np.all(a[:, P:].sum(axis=0) == -2)


def test_l1svm_generator_matches_modeling_layer():
    from pysparselp_b200 import generators
    from pysparselp_b200.examples import example_l1_svm as ex

    x, classes = ex.make_data()
    svm = ex.L1SVM()
    svm.set_data(x, classes)
    c, a_eq, beq, a_in, b_lo, b_up, lb, ub = solver_args_from_lp(svm)
    g, _ = generators.l1svm_lp(1000, 2)
    for u, v in ((g.c, c), (g.lb, lb), (g.ub, ub), (g.b_lower, b_lo), (g.b_upper, b_up),
                 (g.a_ineq.indptr, a_in.indptr), (g.a_ineq.indices, a_in.indices), (g.a_ineq.data, a_in.data)):
        assert np.array_equal(u, v)


def test_random_lp_generator_is_feasible_and_regular():
    from pysparselp_b200 import generators

    lp, xf = generators.random_sparse_lp(500, 900, n_eq=60, nnz_per_row=8, seed=3)
    assert np.all(np.diff(lp.a_ineq.indptr) == 8) and np.all(np.diff(lp.a_eq.indptr) == 8)
    assert np.all(lp.a_ineq @ xf <= lp.b_upper + 1e-12) and np.allclose(lp.a_eq @ xf, lp.b_eq)
    assert np.all(lp.lb <= xf) and np.all(xf <= lp.ub)
    lp2, _ = generators.random_sparse_lp(500, 900, n_eq=60, nnz_per_row=8, seed=3)
    assert np.array_equal(lp.a_ineq.data, lp2.a_ineq.data)


def test_multi_gpu_front_ships_the_lp_through_shared_memory():
    """pysparselp_b200/multi_gpu.py (``n_gpus=``): the LP reaches the helper processes as POSIX shared memory blocks
    described by a small picklable spec; what a helper unpacks is array-identical, CSR matrices included."""
    import pickle

    import numpy as np

    from pysparselp_b200 import generators
    from pysparselp_b200.multi_gpu import pack_lp, unpack_lp

    lp, _ = generators.random_sparse_lp(300, 500, n_eq=40, seed=5)
    args = dict(c=lp.c, a_eq=lp.a_eq, beq=lp.b_eq, a_ineq=lp.a_ineq, b_lower=lp.b_lower, b_upper=lp.b_upper, lb=lp.lb,
                ub=lp.ub, x0=None)
    spec, blocks = pack_lp(args)
    try:
        assert len(pickle.dumps(spec)) < 4096
        got, keep = unpack_lp(pickle.loads(pickle.dumps(spec)))
        for name in ("c", "beq", "b_upper", "lb", "ub"):
            assert np.array_equal(got[name], np.ravel(args[name]))
        assert got["b_lower"] is None and got["x0"] is None
        for name in ("a_eq", "a_ineq"):
            a, b = got[name], args[name]
            assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data) and np.array_equal((a @ np.ones(a.shape[1])), (b @ np.ones(b.shape[1])))
        del got, a, b
        for shm in keep:
            shm.close()
    finally:
        for shm in blocks:
            shm.close()
            shm.unlink()
