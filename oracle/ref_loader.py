"""TEST INFRASTRUCTURE — loader for the *unmodified* reference package.

Imports ``pysparselp`` straight from ``/root/reference`` (read-only, only present
in the build container, never on the GPU box) so that

* the numpy restatement in ``oracle/cpppd_oracle.py`` can be validated against the
  reference's own ``chambolle_pock_ppd`` (``pysparselp/ChambollePockPPD.py:36-346``),
* golden vectors can be minted by ``oracle/make_golden.py``.

Nothing in the product package imports this file.  The shims below only paper
over interpreter/library drift (python 3.12 / numpy 2 vs the reference's pinned
python 3.7 / numpy 1.18) and absent optional dependencies; none touches the
arithmetic of the CP-PPD path:

* ``time.clock``  -> ``time.perf_counter``  (removed in python 3.8; used at
  ``ChambollePockPPD.py:67,244`` and ``SparseLP.py:1016,1378``)
* ``np.float``    -> ``float``              (removed in numpy 1.24; ``SparseLP.py:108,170``)
* ``pysparselp.gaussSiedel`` (Cython, ADMM only, ``ADMM.py:34-35``) -> stub module
* ``matplotlib.pyplot`` -> stub; ``maxflow`` (PyMaxflow) -> scipy ``maximum_flow`` stand-in
"""
import importlib
import os
import sys
import time
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PYSPARSELP_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pysparselp", "ChambollePockPPD.py"))


class _Anything:
    """Swallows any attribute access / call (matplotlib stand-in)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _install_matplotlib_stub():
    try:
        import matplotlib.pyplot  # noqa: F401

        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    anything = _Anything()
    plt.__getattr__ = lambda name: anything  # type: ignore[attr-defined]
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


class _GridGraph:
    """Minimal PyMaxflow ``Graph[int]`` stand-in for grid Potts problems.

    Implements exactly the calls made by ``example_pott_segmentation.py:68-79``
    (add_grid_nodes / add_grid_edges / add_grid_tedges / maxflow / get_grid_segments)
    on top of ``scipy.sparse.csgraph.maximum_flow``.  PyMaxflow labels a node as
    sink-side (True) when it is NOT reachable from the source in the residual
    graph... except that Boykov-Kolmogorov leaves "free" nodes on the source side;
    the equivalent statement is: segment = True iff the node can reach the sink
    in the residual graph.
    """

    def __init__(self, *a):
        self._shape = None
        self._edges = []
        self._t_src = None
        self._t_snk = None

    def add_grid_nodes(self, shape):
        self._shape = tuple(shape)
        return np.arange(int(np.prod(shape))).reshape(shape)

    def add_grid_edges(self, nodeids, weight):
        # default von-Neumann structure, symmetric capacities, all axes
        for axis in range(nodeids.ndim):
            a = np.moveaxis(nodeids, axis, 0)
            self._edges.append((a[:-1].ravel(), a[1:].ravel(), int(weight)))

    def add_grid_tedges(self, nodeids, sourcecaps, sinkcaps):
        self._t_src = np.asarray(sourcecaps).ravel().astype(np.int64)
        self._t_snk = np.asarray(sinkcaps).ravel().astype(np.int64)
        self._ids = nodeids

    def maxflow(self):
        import scipy.sparse as sp
        from scipy.sparse.csgraph import maximum_flow

        nn = int(np.prod(self._shape))
        s, t = nn, nn + 1
        # PyMaxflow cancels negative terminal capacities: tedge(i, a, b) with
        # possibly negative values is equivalent to shifting both so min is 0.
        src = self._t_src.copy()
        snk = self._t_snk.copy()
        mn = np.minimum(src, snk)
        src = src - mn
        snk = snk - mn
        rows, cols, caps = [], [], []
        for u, v, w in self._edges:
            rows += [u, v]
            cols += [v, u]
            caps += [np.full(u.size, w), np.full(u.size, w)]
        ids = np.arange(nn)
        rows += [np.full(nn, s), ids]
        cols += [ids, np.full(nn, t)]
        caps += [src, snk]
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        caps = np.concatenate(caps).astype(np.int32)
        g = sp.csr_matrix((caps, (rows, cols)), shape=(nn + 2, nn + 2))
        res = maximum_flow(g, s, t)
        flow = res.flow
        residual = (g - flow).tocsr()
        residual.data = np.maximum(residual.data, 0)
        residual.eliminate_zeros()
        # nodes that can reach t in the residual graph == BFS from t on the transpose
        from scipy.sparse.csgraph import breadth_first_order

        order = breadth_first_order(residual.T.tocsr(), t, directed=True, return_predecessors=False)
        reach = np.zeros(nn + 2, dtype=bool)
        reach[order] = True
        self._sink_side = reach[:nn]
        return res.flow_value

    def get_grid_segments(self, nodeids):
        return self._sink_side[np.asarray(nodeids)]


class _GraphFactory:
    def __getitem__(self, t):
        return _GridGraph


def _install_maxflow_stub():
    try:
        import maxflow  # noqa: F401

        return
    except Exception:
        pass
    mf = types.ModuleType("maxflow")
    mf.Graph = _GraphFactory()
    sys.modules["maxflow"] = mf


def load_reference():
    """Return the reference ``pysparselp`` package (imported from REFERENCE_ROOT)."""
    if "pysparselp" in sys.modules and getattr(sys.modules["pysparselp"], "_is_reference", False):
        return sys.modules["pysparselp"]
    if not reference_available():
        raise RuntimeError(
            "reference tree not found at %s (it only exists in the build container)" % REFERENCE_ROOT
        )
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter  # type: ignore[attr-defined]
    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]
    _install_matplotlib_stub()
    _install_maxflow_stub()
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    stub = types.ModuleType("pysparselp.gaussSiedel")

    def _unavailable(*a, **k):
        raise RuntimeError("gaussSiedel is stubbed in the oracle loader (ADMM is out of scope)")

    stub.GaussSeidel = _unavailable
    stub.boundedGaussSeidelClass = _unavailable
    sys.modules["pysparselp.gaussSiedel"] = stub
    pkg = importlib.import_module("pysparselp")
    pkg._is_reference = True
    importlib.import_module("pysparselp.SparseLP")
    return pkg


def reference_chambolle_pock_ppd():
    load_reference()
    return importlib.import_module("pysparselp.ChambollePockPPD").chambolle_pock_ppd


def reference_sparse_lp():
    load_reference()
    return importlib.import_module("pysparselp.SparseLP")
