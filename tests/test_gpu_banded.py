"""GPU: banded operands (CPPPD_FLAG_BANDED, csrc/cpppd_banded.cuh) — window-major storage, one launch per window of
the gathered vector, fp64 carries between windows.  The iterates must be the golden bits whatever the window size;
on a random sparse LP of the configs[3] family (randomLP.py:14-75) the banded path must be chosen on its own and
agree bit for bit with the plain-C oracle port and with the SELL kernels."""
import hashlib
import os

import numpy as np
import pytest

from conftest import CASE_PARAMS, GOLDEN_CASES, case_args

pytestmark = pytest.mark.gpu

RANDOM_N = int(os.environ.get("CPPPD_BANDED_RANDOM_N", "1500000"))  # (scaled down to dry-run on the CPU emulation)
WINDOW_MB = os.environ.get("CPPPD_BANDED_TEST_WINDOW_MB", "1")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("window", [7, 1 << 20])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_banded_iterates_vs_golden(name, window, monkeypatch):
    from pysparselp_b200 import _cabi
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    if name == "l1svm":
        window = 199 if window == 7 else 251  # (weight columns of ~1 350 entries: at most 255 per window, one count byte)
    elif window == 7 and name == "potts50":
        window = 997  # (1400 launches per half-iteration otherwise)
    monkeypatch.setenv("CPPPD_BAND_WINDOW", str(window))
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    trace = []
    x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=_cabi.FLAG_BANDED, return_solver=True,
                                         callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
    try:
        info = solver.info()
        y = solver.get_y()
    finally:
        solver.close()
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    assert info["band_in_use"][1] == 1
    if name in ("sc105", "random_small", "afiro", "kb2"):
        assert info["band_in_use"][0] == 1
    if "alpha" not in kw:
        assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold)
    else:
        assert np.max(np.abs(x - g["x_100"])) <= 1e-9 * np.max(np.abs(g["x_100"]))
    got, want = np.array(trace), g["trace_10"]
    assert got.shape == want.shape
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got[fin], want[fin], rtol=1e-6, atol=1e-6 * max(np.max(np.abs(want[fin])), 1e-30))


def test_random_lp_takes_the_banded_path_and_matches_the_c_port(monkeypatch):
    """1.5 M variables, 2.7 M inequalities + 0.3 M equalities, 24 M entries, windows of 1 MB (12 / 23 windows): the
    operands qualify by themselves (no locality, vector > 2 windows), equality and inequality duals are carried
    apart, and x, y equal the C port's bits after 10 iterations — as do the SELL kernels' (CPPPD_FLAG_NO_BANDED)."""
    from oracle.c_port import COracle
    from pysparselp_b200 import _cabi, generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    monkeypatch.setenv("CPPPD_BAND_WINDOW_MB", WINDOW_MB)
    n, m_eq = RANDOM_N, RANDOM_N // 5
    lp, _ = generators.random_sparse_lp(n, 2 * n - m_eq, n_eq=m_eq, seed=3)
    args = generators.lp_args(lp)
    iters = 10
    co = COracle(*args)
    co.iterate(iters)
    want = digest(co.x, co.y)
    results = {}
    for label, flags in (("auto", 0), ("forced", _cabi.FLAG_BANDED), ("sell", _cabi.FLAG_NO_BANDED)):
        x, _, solver = chambolle_pock_ppd(*args, nb_max_iter=iters, nb_iter_plot=iters, flags=flags, return_solver=True)
        try:
            results[label] = (digest(x, solver.get_y()), solver.info())
        finally:
            solver.close()
    assert results["forced"][1]["band_in_use"] == [1, 1]
    assert results["forced"][1]["band_windows"][0] >= 8 and results["forced"][1]["band_windows"][1] >= 16
    assert results["sell"][1]["band_windows"] == [0, 0]
    auto = results["auto"][1]
    assert auto["band_windows"][0] > 0 and auto["band_windows"][1] > 0  # built on its own: sampled locality ~ 1 sector / gather
    assert min(auto["band_sectors_per_gather"]) > 0.6
    for label in results:
        assert results[label][0] == want, label


@pytest.mark.parametrize("name,window", [("random_small", 11), ("l1svm", 251)])
@pytest.mark.parametrize("shape", range(8))
def test_banded_kernel_shapes_give_the_same_bits(shape, name, window, monkeypatch):
    """Every compiled shape of the window kernels — entries through registers (0-3) or staged in shared memory by
    cp.async.bulk + mbarrier (4-7; the L1-SVM tiles take dozens of 384-entry pieces) — produces the golden bits."""
    from pysparselp_b200 import _cabi
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    monkeypatch.setenv("CPPPD_BAND_WINDOW", str(window))
    monkeypatch.setenv("CPPPD_BAND_SHAPE", str(shape))
    args, g = case_args(name)
    x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=100, nb_iter_plot=10, flags=_cabi.FLAG_BANDED, return_solver=True)
    try:
        info = solver.info()
        y = solver.get_y()
    finally:
        solver.close()
    assert info["band_in_use"][1] == 1 and info["band_shape"][1] == shape
    assert np.array_equal(x, g["x_100"])
    assert np.array_equal(y, np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g]))
