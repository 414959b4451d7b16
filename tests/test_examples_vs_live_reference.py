"""CPU, build container only: the reference's OTHER example programs (bipartite matching, k-medians, sparse inverse
covariance — the callers on the modeling side of the path) executed with this package's ``SparseLP`` substituted for
the reference's class: the LP each example hands to ``solve()`` must be array-identical to the one it builds on the
reference class, and ``chambolle_pock_ppd`` on it (CUDA library on the CPU emulator) must return the bits of the
unmodified reference solver.  Skipped where /root/reference does not exist."""
import contextlib
import importlib
import io

import numpy as np
import pytest

from conftest import solver_args_from_lp
from oracle import ref_loader
from oracle.make_golden import lp_digest

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


class _Captured(Exception):
    pass


def lp_handed_to_solve(example, entry, cls, subclass_of=None):
    """Run ``pysparselp.examples.<example>.<entry>()`` with ``cls`` in place of the reference's SparseLP and return
    the model at the moment the example calls ``solve()``."""
    ref_loader.load_reference()
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; example_kmedians.py:74 still uses it
    mod = importlib.import_module("pysparselp.examples." + example)
    holder = {}

    def solve(self, *a, **k):
        holder["lp"] = self
        raise _Captured()

    name = subclass_of or "SparseLP"
    saved = getattr(mod, name)
    extra = {k: v for k, v in vars(saved).items() if callable(v) and not k.startswith("__")} if subclass_of else {}
    setattr(mod, name, type("Recording", (cls,), dict(extra, solve=solve)))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            getattr(mod, entry)()
    except _Captured:
        pass
    finally:
        setattr(mod, name, saved)
    return holder["lp"]


CASES = {
    "bipartite_matching": ("example_bipartite_matching", "run", None),
    "kmedians": ("example_kmedians", "run", None),
    "sparse_inv_covariance": ("example_sparse_inv_covariance", "run", "SparseInvCov"),
}


@pytest.mark.parametrize("case", list(CASES))
def test_reference_examples_run_on_this_modeling_layer(case, monkeypatch):
    import pysparselp_b200.ChambollePockPPD as front
    from emul.patch_plugin import _Adapter
    from pysparselp_b200.SparseLP import SparseLP as Mine

    if case == "sparse_inv_covariance":
        pytest.importorskip("sklearn")
    example, entry, sub = CASES[case]
    theirs = lp_handed_to_solve(example, entry, ref_loader.reference_sparse_lp().SparseLP, sub)
    mine = lp_handed_to_solve(example, entry, Mine, sub)
    args_r, args_m = solver_args_from_lp(theirs), solver_args_from_lp(mine)
    assert lp_digest(args_r) == lp_digest(args_m)
    # the solver on that LP: unmodified reference vs this package over the emulated library
    monkeypatch.setattr(front, "CpPpdSolver", _Adapter)
    kw = dict(nb_max_iter=40, nb_iter_plot=10)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        x_r, _ = ref_loader.reference_chambolle_pock_ppd()(*args_r, **kw)
    x_m, _ = front.chambolle_pock_ppd(*args_m, **kw)
    assert np.array_equal(x_r, x_m, equal_nan=True)
