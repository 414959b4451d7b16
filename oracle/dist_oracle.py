"""TEST INFRASTRUCTURE — CPU (numpy + torch.distributed/gloo) model of the multi-GPU algorithm.

Restates, rank by rank, what libcpppd does with world_size > 1 (csrc/cpppd.cu: setup / exchange):
partition by locality buckets (oracle/partition_oracle.py), owner-computes rows and columns, ghost
copies of xbar / y refreshed by one send/recv per neighbour after each half iteration.  Because every
row sum and column sum stays on one rank and keeps the reference's accumulation order, the result must
be bit-identical to the single-process oracle — which is what `--check` asserts.

    MASTER_ADDR=127.0.0.1 python -m torch.distributed.run --nproc-per-node 2 oracle/dist_oracle.py --check
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import partition_oracle as po  # noqa: E402
from oracle.cpppd_oracle import CpPpdOracle, one_sided_system  # noqa: E402


def _take_rows_keep_order(a, rows, colmap, ncols):
    """Rows `rows` of CSR `a`, entry order untouched, columns renumbered through `colmap`."""
    a = sp.csr_matrix(a)
    lens = np.diff(a.indptr)[rows]
    indptr = np.concatenate(([0], np.cumsum(lens)))
    take = np.concatenate([np.arange(a.indptr[r], a.indptr[r + 1]) for r in rows]) if len(rows) else np.zeros(0, int)
    m = sp.csr_matrix((len(rows), ncols))
    m.data, m.indices, m.indptr = a.data[take], colmap[a.indices[take]].astype(np.int32), indptr.astype(np.int32)
    return m


class RankState:
    def __init__(self, c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, rank, world, theta=1, granule=32):
        if a_eq is not None and a_eq.shape[0] == 0:
            a_eq, beq = None, None
        a_ineq, b_ineq = one_sided_system(a_ineq, b_lower, b_upper)
        blocks = [a for a in (a_eq, a_ineq) if a is not None]
        A = sp.vstack(blocks).tocsr() if len(blocks) > 1 else sp.csr_matrix(blocks[0])
        # vstack sorts nothing but may coalesce formats; rebuild from the blocks to keep entry order
        if len(blocks) > 1:
            A = sp.csr_matrix((np.concatenate([sp.csr_matrix(b).data for b in blocks]),
                               np.concatenate([sp.csr_matrix(b).indices for b in blocks]),
                               np.concatenate(([0], np.cumsum(np.concatenate([np.diff(sp.csr_matrix(b).indptr) for b in blocks]))))),
                              shape=A.shape)
        b = np.concatenate([np.ravel(v) for v in (beq, b_ineq) if v is not None])
        m, n = A.shape
        m_eq = a_eq.shape[0] if a_eq is not None else 0
        self.has_eq, self.has_ineq = a_eq is not None, a_ineq is not None
        part = po.partition(A.indptr, A.indices, n, m_eq, world, granule=granule)
        self.part, self.rank, self.world, self.theta = part, rank, world, theta
        rs, re = part["row_start"][rank], part["row_start"][rank + 1]
        cs, ce = part["col_start"][rank], part["col_start"][rank + 1]
        self.own_rows, self.own_cols = part["row_order"][rs:re], part["col_order"][cs:ce]
        self.ghost_cols, self.ghost_rows = po.ghosts(A.indptr, A.indices, part, rank)
        self.nloc, self.mloc = self.own_cols.size, self.own_rows.size
        colmap = np.full(n, -1)
        colmap[np.concatenate((self.own_cols, self.ghost_cols))] = np.arange(self.nloc + self.ghost_cols.size)
        rowmap = np.full(m, -1)
        rowmap[np.concatenate((self.own_rows, self.ghost_rows))] = np.arange(self.mloc + self.ghost_rows.size)
        self.A = _take_rows_keep_order(A, self.own_rows, colmap, self.nloc + self.ghost_cols.size)
        # columns as rows of A^T, entries in original row order, equality and inequality parts apart
        At = A.tocsc()  # csr -> csc walks rows in order: entries of a column sorted by original row
        At = sp.csr_matrix((At.data, At.indices, At.indptr), shape=(n, m))
        eq_mask = At.indices < m_eq
        def part_of(mask):
            """Own columns as rows of A^T restricted to the entries selected by `mask` (order kept)."""
            rows = _take_rows_keep_order(At, self.own_cols, rowmap, self.mloc + self.ghost_rows.size)
            sel = _take_mask(At, self.own_cols, mask)
            row_of_entry = np.repeat(np.arange(rows.shape[0]), np.diff(rows.indptr))
            out = sp.csr_matrix(rows.shape)
            out.data, out.indices = rows.data[sel], rows.indices[sel]
            out.indptr = np.concatenate(
                ([0], np.cumsum(np.bincount(row_of_entry[sel], minlength=rows.shape[0])))).astype(np.int32)
            return out

        self.At_eq, self.At_in = part_of(eq_mask), part_of(~eq_mask)
        self.m_eq_loc = int(np.sum(self.own_rows < m_eq))
        self.c, self.lb, self.ub = c[self.own_cols], lb[self.own_cols], ub[self.own_cols]
        self.b = b[self.own_rows]
        ref = CpPpdOracle(c, a_eq, beq, a_ineq, None, b_ineq, lb, ub)  # preconditioners: complete rows/cols are local
        self.T = ref.diag_t[self.own_cols]
        sig = np.concatenate([v for v in (ref.sig_eq, ref.sig_ineq) if v is not None])
        self.sigma = sig[self.own_rows]
        self.x = np.zeros(self.nloc)
        self.xbar = np.zeros(self.nloc + self.ghost_cols.size)
        self.y = np.zeros(self.mloc + self.ghost_rows.size)
        # who sends what: ghosts grouped by owner (ascending position == exchange order)
        self.recv_x = self._by_owner(self.ghost_cols, part["col_owner"])
        self.recv_y = self._by_owner(self.ghost_rows, part["row_owner"])
        self.n, self.m = n, m

    @staticmethod
    def _by_owner(ids, owner):
        out, at = {}, 0
        for o in np.unique(owner[ids]) if ids.size else []:
            cnt = int(np.sum(owner[ids] == o))
            out[int(o)] = (at, cnt)
            at += cnt
        return out

    def send_lists(self, all_ghosts, local_of):
        """What this rank sends to every peer, from the peers' ghost lists (all ranks know the pattern)."""
        out = {}
        for t, ids in all_ghosts.items():
            mine = ids[local_of[ids] >= 0]
            if t != self.rank and mine.size:
                out[t] = local_of[mine]
        return out

    def primal(self):
        d = self.c
        if self.has_eq:
            d = d + self.At_eq @ self.y
        if self.has_ineq:
            d = d + self.At_in @ self.y
        x2 = self.x - self.T * d
        np.maximum(x2, self.lb, x2)
        np.minimum(x2, self.ub, x2)
        self.xbar[: self.nloc] = (1 + self.theta) * x2 - self.theta * self.x
        self.x = x2

    def dual(self):
        r = self.A @ self.xbar - self.b
        yn = self.y[: self.mloc] + self.sigma * r
        yn[self.m_eq_loc:] = np.maximum(yn[self.m_eq_loc:], 0)
        self.y[: self.mloc] = yn


def _take_mask(At, rows, mask):
    return np.concatenate([mask[At.indptr[r]: At.indptr[r + 1]] for r in rows]) if len(rows) else np.zeros(0, bool)


def exchange(dist, torch, vec, owned, recv, send):
    reqs, bufs = [], {}
    for t, idx in sorted(send.items()):
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(vec[idx])), dst=t))
    for o, (at, cnt) in sorted(recv.items()):
        bufs[o] = torch.empty(cnt, dtype=torch.float64)
        reqs.append(dist.irecv(bufs[o], src=o))
    for r in reqs:
        r.wait()
    for o, (at, cnt) in recv.items():
        vec[owned + at: owned + at + cnt] = bufs[o].numpy()


def run_distributed(args, iters, granule=32):
    """Every rank: iterate `iters` times, return the assembled full (x, y)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    st = RankState(*args, rank=rank, world=world, granule=granule)
    # every rank derives every rank's ghost lists from the pattern (as the CUDA setup does)
    c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub = args
    states_ghost_cols, states_ghost_rows = {}, {}
    blocks = [sp.csr_matrix(a) for a in (a_eq if a_eq is not None and a_eq.shape[0] else None,
                                         one_sided_system(a_ineq, b_lower, b_upper)[0]) if a is not None]
    indptr = np.concatenate(([0], np.cumsum(np.concatenate([np.diff(b_.indptr) for b_ in blocks]))))
    indices = np.concatenate([b_.indices for b_ in blocks])
    for t in range(world):
        states_ghost_cols[t], states_ghost_rows[t] = po.ghosts(indptr, indices, st.part, t)
    col_local = np.full(st.n, -1)
    col_local[st.own_cols] = np.arange(st.nloc)
    row_local = np.full(st.m, -1)
    row_local[st.own_rows] = np.arange(st.mloc)
    send_x = st.send_lists(states_ghost_cols, col_local)
    send_y = st.send_lists(states_ghost_rows, row_local)
    for _ in range(iters):
        st.primal()
        exchange(dist, torch, st.xbar, st.nloc, st.recv_x, send_x)
        st.dual()
        exchange(dist, torch, st.y, st.mloc, st.recv_y, send_y)
    full_x = torch.zeros(st.n, dtype=torch.float64)
    full_x[torch.from_numpy(st.own_cols.astype(np.int64))] = torch.from_numpy(st.x)
    full_y = torch.zeros(st.m, dtype=torch.float64)
    full_y[torch.from_numpy(st.own_rows.astype(np.int64))] = torch.from_numpy(st.y[: st.mloc])
    dist.all_reduce(full_x)
    dist.all_reduce(full_y)
    return full_x.numpy(), full_y.numpy(), st


def _check():
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import case_args

    dist.init_process_group("gloo")
    for name in ("potts50", "sc105", "random_small"):
        args, g = case_args(name)
        x, y, st = run_distributed(args, 100)
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        assert np.array_equal(x, g["x_100"]), name
        assert np.array_equal(y, y_gold), name
        if dist.get_rank() == 0:
            print("dist oracle %s ok: world %d, ghosts cols %d rows %d" % (
                name, dist.get_world_size(), st.ghost_cols.size, st.ghost_rows.size), flush=True)
    dist.barrier()
    if dist.get_rank() == 0:
        print("DIST_ORACLE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    if "--check" in sys.argv:
        _check()
