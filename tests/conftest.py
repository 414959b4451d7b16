import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a GPU: they are skipped unless selected with -m gpu."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or os.environ.get("CPPPD_EMULATE_GPU_TESTS"):  # see tests/emul/patch_plugin.py
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def unpack_lp(g):
    """Solver inputs stored inside a golden file by oracle/make_golden.pack_lp."""
    def mat(tag):
        key = "lp_a%s_data" % tag
        if key not in g:
            return None
        return sp.csr_matrix((g[key], g["lp_a%s_indices" % tag], g["lp_a%s_indptr" % tag]),
                             shape=tuple(g["lp_a%s_shape" % tag]))
    return (g["lp_c"], mat("eq"), g.get("lp_beq"), mat("ineq"), g.get("lp_b_lower"), g.get("lp_b_upper"),
            g["lp_lb"], g["lp_ub"])


def solver_args_from_lp(lp):
    """What SparseLP.solve hands to chambolle_pock_ppd (reference SparseLP.py:1244-1287)."""
    import copy

    red = copy.deepcopy(lp)
    red.remove_fixed_variables()
    return (red.costsvector, red.a_equalities, red.b_equalities, red.a_inequalities, red.b_lower,
            red.b_upper, red.lower_bounds, red.upper_bounds)


def case_args(name):
    """Rebuild the solver inputs of a golden case WITHOUT the reference tree and check their digest."""
    from oracle.make_golden import lp_digest

    g = load_golden(name)
    if name == "potts50":
        from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

        lp, _, _, _ = build_linear_program(50, 0.5, 500, with_ground_truth=False)
        args = solver_args_from_lp(lp)
    elif name == "l1svm":
        from pysparselp_b200.examples import example_l1_svm as ex

        x, classes = ex.make_data()
        svm = ex.L1SVM()
        svm.set_data(x, classes)
        args = solver_args_from_lp(svm)
    else:
        args = unpack_lp(g)
    assert lp_digest(args) == bytes(g["digest"]).decode(), "rebuilt LP differs from the one the golden was minted on"
    return args, g


# potts50 / sc105 / l1svm: the LPs of the reference's own regression tests; afiro, kb2 (the one with a BOUNDS section),
# sc50a, sc50b: the other netlib problems the reference vendors, through MPS parser -> modeling layer -> solver
GOLDEN_CASES = ["potts50", "sc105", "l1svm", "random_small", "random_small_alpha", "afiro", "kb2", "sc50a", "sc50b"]
CASE_PARAMS = {"random_small_alpha": dict(alpha=1.5, theta=0.7)}
