"""CPU coverage of the N > 1 path: the partition contract and a world_size-2 gloo run of the
distributed algorithm model (oracle/dist_oracle.py), which must reproduce the reference goldens
bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT
from oracle import partition_oracle as po


def _stacked(lp):
    blocks = [a for a in (lp.a_eq, lp.a_ineq) if a is not None]
    a = sp.vstack(blocks).tocsr() if len(blocks) > 1 else blocks[0]
    return a, (lp.a_eq.shape[0] if lp.a_eq is not None else 0)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_is_a_valid_cover(world):
    from pysparselp_b200 import generators

    for lp in (generators.potts_lp(24), generators.random_sparse_lp(300, 500, n_eq=80, seed=2)[0]):
        a, m_eq = _stacked(lp)
        part = po.partition(a.indptr, a.indices, a.shape[1], m_eq, world, granule=32)
        assert np.array_equal(np.sort(part["row_order"]), np.arange(a.shape[0]))
        assert np.array_equal(np.sort(part["col_order"]), np.arange(a.shape[1]))
        assert part["row_start"][-1] == a.shape[0] and part["col_start"][-1] == a.shape[1]
        for r in range(world):
            rows = part["row_order"][part["row_start"][r]: part["row_start"][r + 1]]
            k = part["m_eq_local"][r]
            assert np.all(rows[:k] < m_eq) and np.all(rows[k:] >= m_eq)  # equalities first
            assert np.all(part["row_owner"][rows] == r)
        # same inputs -> same partition (pure function)
        again = po.partition(a.indptr, a.indices, a.shape[1], m_eq, world, granule=32)
        assert all(np.array_equal(part[k], again[k]) for k in ("row_order", "col_order", "row_start", "col_start"))


def test_potts_partition_has_thin_halos():
    from pysparselp_b200 import generators

    lp = generators.potts_lp(96)
    a, m_eq = _stacked(lp)
    part = po.partition(a.indptr, a.indices, a.shape[1], m_eq, 4, granule=32)
    work = np.diff(part["row_start"]) + np.diff(part["col_start"])
    assert work.max() / work.min() < 1.1  # balanced strips
    for r in range(4):
        gc, gr = po.ghosts(a.indptr, a.indices, part, r)
        assert gc.size <= 96 and gr.size <= 4 * 96  # one image row of pixels / its four edge-row blocks
        # ghosts are exactly the foreign columns / rows the rank touches
        rows = part["row_order"][part["row_start"][r]: part["row_start"][r + 1]]
        touched = np.unique(a[rows].indices)
        assert np.array_equal(np.sort(gc), touched[part["col_owner"][touched] != r])


def test_world2_gloo_matches_reference_goldens():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "oracle", "dist_oracle.py"), "--check"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0 and "DIST_ORACLE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
