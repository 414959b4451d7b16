"""Direct CSR emitters for the large synthetic LPs of BASELINE.json (host side, numpy).

The modeling layer is convenient but builds matrices through several temporaries; at
benchmark sizes (4096x4096 Potts: 201 M stored entries) the LP is emitted here straight
into its final CSR arrays, in exactly the row / column / entry order the modeling layer
produces (checked for array equality at small sizes in ``tests/test_generators.py``).

Every generator returns an ``LPArrays`` — the argument list of ``chambolle_pock_ppd``.
"""
from collections import namedtuple

import numpy as np
import scipy.sparse as sp

LPArrays = namedtuple("LPArrays", "c a_eq b_eq a_ineq b_lower b_upper lb ub")


def lp_args(lp):
    """Positional arguments for ``chambolle_pock_ppd`` from an ``LPArrays``."""
    return (lp.c, lp.a_eq, lp.b_eq, lp.a_ineq, lp.b_lower, lp.b_upper, lp.lb, lp.ub)


def _csr_unchecked(data, indices, indptr, shape):
    m = sp.csr_matrix(shape, dtype=np.float64)
    m.data, m.indices, m.indptr = data, indices, indptr
    return m


def potts_lp(height, width=None, coef_potts=0.5, coef_mul=500, seed=1, empty=np.empty):
    """Potts segmentation LP of ``examples/example_pott_segmentation.build_linear_program``.

    Reference generator: ``pysparselp/examples/example_pott_segmentation.py:15-51, :54-92``.
    Variables: pixels ``i*W + j`` (cost unary/coef_mul, bounds [0,1]); horizontal-edge
    auxiliaries ``P + i*(W-1) + j``; vertical-edge auxiliaries ``P + H*(W-1) + i*W + j``
    (cost round(coef_potts*coef_mul)/coef_mul, bounds [0,1]).  Rows, 3 entries each, all
    ``<= 0``: h(+), h(-), v(+), v(-) blocks in raster order of the edge; columns of a row are
    ``(second pixel, first pixel, aux)`` with values ``(1,-1,-1)`` then ``(-1,1,-1)``.
    ``empty`` lets the caller provide the array allocator (e.g. pinned host memory).
    """
    H = int(height)
    W = H if width is None else int(width)
    rs = np.random.RandomState(seed)
    unary = np.round(coef_mul * (rs.rand(H, W, 1) * 2 - 1))
    pair_cost = round(coef_potts * coef_mul) / coef_mul
    P = H * W
    n_h, n_v = H * (W - 1), (H - 1) * W
    n = P + n_h + n_v
    m = 2 * (n_h + n_v)
    c = empty(n, dtype=np.float64)
    c[:P] = (unary / coef_mul).ravel()
    c[P:] = pair_cost
    lb = empty(n, dtype=np.float64)
    lb[:] = 0.0
    ub = empty(n, dtype=np.float64)
    ub[:] = 1.0
    idx_dtype = np.int32 if max(3 * m, n) < 2**31 - 1 else np.int64
    pix = np.arange(P, dtype=idx_dtype).reshape(H, W)
    indices = empty(3 * m, dtype=idx_dtype).reshape(m, 3)
    data = empty(3 * m, dtype=np.float64).reshape(m, 3)
    row = 0
    for first, second, aux0, count in (
        (pix[:, :-1].ravel(), pix[:, 1:].ravel(), P, n_h),
        (pix[:-1, :].ravel(), pix[1:, :].ravel(), P + n_h, n_v),
    ):
        aux = np.arange(aux0, aux0 + count, dtype=idx_dtype)
        for sign in (1.0, -1.0):
            blk = slice(row, row + count)
            indices[blk, 0] = second
            indices[blk, 1] = first
            indices[blk, 2] = aux
            data[blk, 0] = sign
            data[blk, 1] = -sign
            data[blk, 2] = -1.0
            row += count
    indptr = empty(m + 1, dtype=idx_dtype)
    indptr[:] = np.arange(0, 3 * m + 1, 3, dtype=idx_dtype)
    a_ineq = _csr_unchecked(data.reshape(-1), indices.reshape(-1), indptr, (m, n))
    b_upper = empty(m, dtype=np.float64)
    b_upper[:] = 0.0
    return LPArrays(c, None, None, a_ineq, None, b_upper, lb, ub)


def random_sparse_lp(nbvar, n_ineq, n_eq=0, nnz_per_row=8, seed=0):
    """Sparse restatement of the reference's ``randomLP.generate_random_lp``.

    Reference: ``pysparselp/randomLP.py:14-75`` — which draws *dense* ``randn(n_ineq, nbvar)``
    masked by ``rand < sparsity`` and therefore cannot scale (and does not import as
    shipped).  This restatement keeps its distributions — values, costs, feasible point and
    bound offsets are ``round(randn*100)/100``; ``b_upper = ceil((A x_f + |noise|)*1000)/1000``;
    ``b_lower = None``; ``lb = x_f + min(0,t)``, ``ub = x_f + max(0,t)`` — with a pattern of
    exactly ``nnz_per_row`` distinct columns per row (the reference drops rows with fewer than
    two entries, ``:40-42,:64-67``; here every row has ``nnz_per_row >= 2``), drawn from
    ``np.random.default_rng(seed)``.  Rows are feasible at ``x_f`` by construction.
    """
    rng = np.random.default_rng(seed)
    n = int(nbvar)

    def rvals(size):
        return np.round(rng.standard_normal(size) * 100) / 100

    def pattern(rows):
        k = nnz_per_row
        cols = rng.integers(0, n, size=(rows, k), dtype=np.int64)
        # re-draw duplicates inside a row (rare when k << n)
        while True:
            srt = np.sort(cols, axis=1)
            bad = np.flatnonzero(np.any(srt[:, 1:] == srt[:, :-1], axis=1))
            if bad.size == 0:
                break
            cols[bad] = rng.integers(0, n, size=(bad.size, k), dtype=np.int64)
        cols.sort(axis=1)
        vals = rvals((rows, k))
        vals[vals == 0] = 0.01  # a stored zero would not count as an entry in the reference
        indptr = np.arange(0, rows * k + 1, k, dtype=np.int64)
        idt = np.int32 if rows * k < 2**31 - 1 else np.int64
        return _csr_unchecked(vals.reshape(-1), cols.reshape(-1).astype(np.int32), indptr.astype(idt), (rows, n))

    x_f = rvals(n)
    a_ineq = pattern(int(n_ineq))
    b_upper = np.ceil((a_ineq @ x_f + np.abs(rvals(int(n_ineq)))) * 1000) / 1000
    costs = rvals(n)
    t = rvals(n)
    lb = x_f + np.minimum(0, t)
    ub = x_f + np.maximum(0, t)
    a_eq = b_eq = None
    if n_eq > 0:
        a_eq = pattern(int(n_eq))
        b_eq = a_eq @ x_f
    return LPArrays(costs, a_eq, b_eq, a_ineq, None, b_upper, lb, ub), x_f


def random_sparse_lp_chunked(nbvar, n_ineq, n_eq=0, nnz_per_row=8, seed=0, chunk_rows=1 << 20, threads=None, empty=np.empty):
    """``random_sparse_lp`` for benchmark sizes (BASELINE configs[3]: 20 M variables, 40 M rows, 320 M entries).

    Same distributions and the same LP family as ``random_sparse_lp`` (reference ``pysparselp/randomLP.py:14-75``,
    see there), but the rows are drawn in independent chunks of ``chunk_rows`` rows, every chunk from its own child
    of ``np.random.SeedSequence(seed)``, written straight into the final CSR arrays.  The result depends on
    ``(seed, chunk_rows)`` only — not on ``threads``, the number of worker threads the chunks are spread over
    (numpy's generators and sorts release the GIL).  ``empty`` is the array allocator (e.g. pinned host memory).
    """
    from concurrent.futures import ThreadPoolExecutor
    import os

    n, k = int(nbvar), int(nnz_per_row)
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    threads = max(1, min(int(threads), 64))
    root = np.random.SeedSequence(seed)
    vec_seed, eq_seed, ineq_seed = root.spawn(3)

    def rvals(rng, size):
        v = rng.standard_normal(size)
        v *= 100
        np.round(v, out=v)
        v /= 100
        return v

    # vectors of length n, drawn in chunks as well
    ncols_chunks = [(s, min(n, s + chunk_rows)) for s in range(0, n, chunk_rows)]
    x_f, costs, t = (empty(n, dtype=np.float64) for _ in range(3))

    def fill_vec(job):
        (s, e), ss = job
        rng = np.random.default_rng(ss)
        x_f[s:e] = rvals(rng, e - s)
        costs[s:e] = rvals(rng, e - s)
        t[s:e] = rvals(rng, e - s)

    def block(rows, ss_root):
        rows = int(rows)
        indices = empty(rows * k, dtype=np.int32)
        data = empty(rows * k, dtype=np.float64)
        rhs = empty(rows, dtype=np.float64)
        noise = empty(rows, dtype=np.float64)
        chunks = [(s, min(rows, s + chunk_rows)) for s in range(0, rows, chunk_rows)]

        def fill(job):
            (s, e), ss = job
            rng = np.random.default_rng(ss)
            cols = rng.integers(0, n, size=(e - s, k), dtype=np.int32)
            cols.sort(axis=1)
            while True:  # re-draw the (rare) rows with a repeated column
                bad = np.flatnonzero(np.any(cols[:, 1:] == cols[:, :-1], axis=1))
                if bad.size == 0:
                    break
                cols[bad] = np.sort(rng.integers(0, n, size=(bad.size, k), dtype=np.int32), axis=1)
            vals = rvals(rng, (e - s, k))
            vals[vals == 0] = 0.01
            indices[s * k:e * k] = cols.reshape(-1)
            data[s * k:e * k] = vals.reshape(-1)
            acc = np.zeros(e - s)
            for q in range(k):  # A x_f accumulated in stored order, like scipy's csr_matvec
                acc += vals[:, q] * x_f[cols[:, q]]
            rhs[s:e] = acc
            noise[s:e] = np.abs(rvals(rng, e - s))

        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(fill, zip(chunks, ss_root.spawn(len(chunks)))))
        idt = np.int32 if rows * k < 2**31 - 1 else np.int64
        indptr = empty(rows + 1, dtype=idt)
        indptr[:] = np.arange(0, rows * k + 1, k, dtype=idt)
        return _csr_unchecked(data, indices, indptr, (rows, n)), rhs, noise

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(fill_vec, zip(ncols_chunks, vec_seed.spawn(len(ncols_chunks)))))
    a_ineq, rhs, noise = block(n_ineq, ineq_seed)
    b_upper = rhs
    b_upper += noise
    b_upper *= 1000
    np.ceil(b_upper, out=b_upper)
    b_upper /= 1000
    lb = empty(n, dtype=np.float64)
    ub = empty(n, dtype=np.float64)
    np.add(x_f, np.minimum(0, t), out=lb)
    np.add(x_f, np.maximum(0, t), out=ub)
    a_eq = b_eq = None
    if n_eq > 0:
        a_eq, b_eq, _ = block(n_eq, eq_seed)
    return LPArrays(costs, a_eq, b_eq, a_ineq, None, b_upper, lb, ub), x_f


def l1svm_lp(nb_examples, nb_features, nb_classes=3, seed=1):
    """L1-SVM LP of ``examples/example_l1_svm.L1SVM.set_data`` emitted directly.

    Reference: ``pysparselp/examples/example_l1_svm.py:13-68, :95-104``.  Variables: weights
    (K x (F+1), free), abs-penalty auxiliaries (K(F+1), >= 0, cost 1), slacks (N, >= 0, cost 1).
    Rows: ``w - a <= 0`` block, ``-w - a <= 0`` block, then for every class k the rows
    ``W[y_i].xh_i - W[k].xh_i + eps_i >= 1`` over the examples with ``y_i != k``.
    Returns two-sided bounds (``b_lower`` / ``b_upper``) like the modeling layer does.
    """
    rs = np.random.RandomState(seed)
    N, F, K = int(nb_examples), int(nb_features), int(nb_classes)
    x = rs.rand(N, F)
    xh = np.hstack((x, np.ones((N, 1))))
    w = rs.randn(K, F)
    w = w / np.sum(w ** 2, axis=1)[:, None]
    w = np.hstack((w, -0.5 * np.sum(w, axis=1)[:, None]))
    classes = np.argmax(w.dot(xh.T).T, axis=1)
    nw = K * (F + 1)
    n = 2 * nw + N
    widx = np.arange(nw).reshape(K, F + 1)
    aidx = nw + np.arange(nw)
    eidx = 2 * nw + np.arange(N)
    c = np.concatenate((np.zeros(nw), np.ones(nw), np.ones(N)))
    lb = np.concatenate((np.full(nw, -np.inf), np.zeros(nw), np.zeros(N)))
    ub = np.full(n, np.inf)
    blocks, lo, up = [], [], []
    for sign in (1.0, -1.0):
        cols = np.column_stack((widx.ravel(), aidx))
        vals = np.tile(np.array([sign, -1.0]), (nw, 1))
        blocks.append(_rows(cols, vals, n))
        lo.append(np.full(nw, -np.inf))
        up.append(np.zeros(nw))
    own = widx[classes, :]
    for k in range(K):
        keep = classes != k
        cols = np.column_stack((own, np.tile(widx[[k], :], (N, 1)), eidx[:, None]))[keep]
        vals = np.column_stack((xh, -xh, np.ones((N, 1))))[keep]
        blocks.append(_rows(cols, vals, n))
        lo.append(np.ones(int(keep.sum())))
        up.append(np.full(int(keep.sum()), np.inf))
    a = sp.vstack(blocks).tocsr()
    return LPArrays(c, None, None, a, np.concatenate(lo), np.concatenate(up), lb, ub), (x, classes, widx)


def _rows(cols, vals, n):
    """Constant-width rows, zeros dropped, order kept (``SparseLP.crd_matrix`` semantics)."""
    keep = vals != 0
    indptr = np.concatenate(([0], np.cumsum(keep.sum(axis=1))))
    return sp.csr_matrix((vals[keep], cols[keep], indptr), shape=(cols.shape[0], n))
