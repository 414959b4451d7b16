"""TEST INFRASTRUCTURE — numpy restatement of the multi-GPU partition of libcpppd.

The reference has no distributed code; the partition is this build's own contract
("a pure function of (indptr, indices, N), exposed through the ABI, tested bit-exact
against a Python restatement").  The CUDA library computes exactly the same integers on the
device (csrc/cpppd.cu: setup()); ``tests/test_gpu_parity.py::test_partition_matches_python_restatement``
and ``tests/dist_worker.py`` compare them, ``tests/test_dist_cpu.py`` checks its invariants on CPU.

Definitions (A = [A_eq; A_ineq], m x n, CSR; N ranks; granule G):
  row_key[i] = min column index of row i                (n for an empty row)
  col_key[j] = min row_key over the rows that hit j     (n for an empty column)
  bucket     = key // G,   nb = n // G + 2 buckets
  work[q]    = nnz of the rows in bucket q + nnz of the columns in bucket q
  owner(q)   = min(N-1, (work before bucket q) * N // total work)
A row / column belongs to the owner of its bucket — unless that leaves some rank with more than 1.5 times
its share of the row entries or of the column entries (patterns without locality: every row of an L1-SVM LP
starts at a weight column, so all rows share one bucket; a random pattern gives the first ranks the columns
and the last ranks the rows).  Then ownership falls back to the BALANCED SPLIT, decided per row and per
column from the prefix sums in original order:
  owner(row i)    = min(N-1, indptr[i] * N // nnz)
  owner(column j) = min(N-1, (entries of the columns before j) * N // nnz)
Rank r keeps its rows in the order
(equalities first, then inequalities; inside each, by bucket, then by min(length, 4095), then by
original index) and its columns in the order (bucket, min(length, 4095), original index) — grouping
equal lengths inside a bucket is the sigma-sorting of SELL-C-sigma: slices of 32 neighbours get
nearly equal widths.  With banded operands (CPPPD_FLAG_BANDED, or chosen automatically for patterns without
locality) and the balanced split, rows and columns instead keep their ORIGINAL order inside (rank, kind).
Ghost columns of rank r: columns owned by another rank that appear in r's rows; ghost rows:
rows owned by another rank that hit r's columns.  Both are listed by (owner, local position).
"""
import numpy as np


def default_granule(n):
    g = 32
    while g < (n >> 14):
        g *= 2
    return g


def locality_keys(indptr, indices, n):
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    m = indptr.size - 1
    lens = np.diff(indptr)
    row_key = np.full(m, n, dtype=np.int64)
    ne = np.flatnonzero(lens > 0)
    if ne.size:
        row_key[ne] = np.minimum.reduceat(indices, indptr[:-1][ne])
    col_key = np.full(n, n, dtype=np.int64)
    np.minimum.at(col_key, indices, np.repeat(row_key, lens))
    return row_key, col_key


def partition(indptr, indices, n, m_eq, world, granule=None, reorder=True, keep_order=False):
    """Returns dict(row_order, row_start, col_order, col_start, m_eq_local).

    ``row_order[row_start[r]:row_start[r+1]]`` are the original row ids owned by rank r in local
    order; likewise for columns.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    m = indptr.size - 1
    G = default_granule(n) if granule is None else int(granule)
    lens = np.diff(indptr)
    col_lens = np.bincount(indices, minlength=n).astype(np.int64)
    if reorder:
        row_key, col_key = locality_keys(indptr, indices, n)
        rq, cq = row_key // G, col_key // G
    else:
        rq, cq = np.zeros(m, dtype=np.int64), np.zeros(n, dtype=np.int64)
    nb = n // G + 2
    work = (np.bincount(rq, weights=lens, minlength=nb) + np.bincount(cq, weights=col_lens, minlength=nb)).astype(np.int64)
    before = np.cumsum(work) - work
    total = int(work.sum())
    if total > 0:
        owner_of_bucket = np.minimum(world - 1, before * world // total)
    else:
        owner_of_bucket = np.zeros(nb, dtype=np.int64)
    row_owner, col_owner = owner_of_bucket[rq], owner_of_bucket[cq]
    nnz = int(lens.sum())
    row_share = np.bincount(row_owner, weights=lens, minlength=world).astype(np.int64)
    col_share = np.bincount(col_owner, weights=col_lens, minlength=world).astype(np.int64)
    balanced = bool(nnz > 0 and 2 * world * max(int(row_share.max()), int(col_share.max())) > 3 * nnz)
    if balanced:
        row_owner = np.minimum(world - 1, indptr[:-1] * world // nnz)
        col_owner = np.minimum(world - 1, (np.cumsum(col_lens) - col_lens) * world // nnz)
    is_ineq = (np.arange(m) >= m_eq).astype(np.int64)
    if balanced and keep_order:
        # banded operands (csrc/cpppd_banded.cuh) on several GPUs: original order inside (owner, kind), so that a
        # range of original ids stays a few contiguous pieces of the local layout
        row_order = np.lexsort((np.arange(m), is_ineq, row_owner))
        col_order = np.lexsort((np.arange(n), col_owner))
    else:
        row_order = np.lexsort((np.arange(m), np.minimum(lens, 4095), rq, is_ineq, row_owner))
        col_order = np.lexsort((np.arange(n), np.minimum(col_lens, 4095), cq, col_owner))
    row_start = np.concatenate(([0], np.cumsum(np.bincount(row_owner, minlength=world))))
    col_start = np.concatenate(([0], np.cumsum(np.bincount(col_owner, minlength=world))))
    m_eq_local = np.bincount(row_owner[:m_eq], minlength=world)
    return dict(row_order=row_order, row_start=row_start, col_order=col_order, col_start=col_start,
                m_eq_local=m_eq_local, row_owner=row_owner, col_owner=col_owner, granule=G, balanced_split=balanced)


def ghosts(indptr, indices, part, rank):
    """(ghost column ids, ghost row ids) of `rank`, original ids, in exchange order."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    m = indptr.size - 1
    n = part["col_owner"].size
    rows_of_entries = np.repeat(np.arange(m), np.diff(indptr))
    ro = part["row_owner"][rows_of_entries]
    co = part["col_owner"][indices]
    col_pos = np.empty(n, dtype=np.int64)
    col_pos[part["col_order"]] = np.arange(n)
    row_pos = np.empty(m, dtype=np.int64)
    row_pos[part["row_order"]] = np.arange(m)
    gc = np.unique(indices[(ro == rank) & (co != rank)])
    gr = np.unique(rows_of_entries[(co == rank) & (ro != rank)])
    gc = gc[np.argsort(col_pos[gc], kind="stable")]
    gr = gr[np.argsort(row_pos[gr], kind="stable")]
    return gc, gr
