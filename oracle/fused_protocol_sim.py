"""TEST INFRASTRUCTURE — thread-level model of the halo exchange fused into k_primal / k_dual.

One Python thread per rank plays the two kernels of csrc/cpppd_hot_kernels.cuh slice by slice, in a
random slice order and with random stalls, using the same protocol as the CUDA code:

  * role[s] bit 0: the slice holds a row that is sent; bit 1: it reads a ghost entry;
  * a slice with a role first waits until every expected neighbour stamp of the *consumed* halo has
    reached `want` (k_dual(k) wants the xbar stamp k+1, k_primal(k) the y stamp k);
  * sent values are stored straight into the neighbours' ghost slots;
  * when all slices of the "kernel" are done the new stamp of the *produced* halo is published to the
    neighbours and the consumed counter advances.

Ghost vectors and flag arrays are plain numpy arrays shared between the threads (the stand-in for
CUDA IPC peer memory).  A wrong stamp rule or a missing wait shows up as a dead-lock (time-out) or as
iterates that differ from the single-process oracle.
"""
import random
import threading
import time

import numpy as np

from oracle import partition_oracle as po
from oracle.dist_oracle import RankState


class FusedRank:
    def __init__(self, args, rank, world, granule, rng):
        self.st = RankState(*args, rank=rank, world=world, granule=granule)
        self.rank, self.world, self.rng = rank, world, rng
        self.flags = np.zeros(2 * world, dtype=np.int64)   # [kind * world + src] stamps written by the peers
        self.push_stamp = [0, 0]
        self.wait_stamp = [0, 0]

    def link(self, ranks, ghosts_cols, ghosts_rows):
        st, me = self.st, self.rank
        col_local = np.full(st.n, -1)
        col_local[st.own_cols] = np.arange(st.nloc)
        row_local = np.full(st.m, -1)
        row_local[st.own_rows] = np.arange(st.mloc)
        # send tables: (local index) -> list of (peer, destination slot in the peer's vector)
        self.send = [dict(), dict()]
        self.send_peers = [set(), set()]
        for t, other in enumerate(ranks):
            if t == me:
                continue
            for kind, ghosts, local_of, owned_t, recv in (
                    (0, ghosts_cols[t], col_local, other.st.nloc, other.st.recv_x),
                    (1, ghosts_rows[t], row_local, other.st.mloc, other.st.recv_y)):
                if me not in recv:
                    continue
                at, cnt = recv[me]
                mine = ghosts[local_of[ghosts] >= 0]
                assert mine.size == cnt
                for k, gid in enumerate(mine):
                    self.send[kind].setdefault(int(local_of[gid]), []).append((t, owned_t + at + k))
                self.send_peers[kind].add(t)
        self.recv_peers = [set(st.recv_x), set(st.recv_y)]
        self.ranks = ranks
        # slice roles
        def roles(matrix_rows_entries, nrows, owned_other, sent):
            ns = (nrows + 31) // 32
            role = np.zeros(ns, dtype=np.int64)
            for i in sent:
                role[i >> 5] |= 1
            for s in range(ns):
                rows = slice(32 * s, min(32 * s + 32, nrows))
                idx = matrix_rows_entries(rows)
                if idx.size and idx.max() >= owned_other:
                    role[s] |= 2
            return role
        at_idx = lambda rows: np.concatenate((st.At_eq[rows].indices, st.At_in[rows].indices))
        a_idx = lambda rows: st.A[rows].indices
        self.role = [roles(at_idx, st.nloc, st.mloc, self.send[0].keys()),
                     roles(a_idx, st.mloc, st.nloc, self.send[1].keys())]

    # ---- protocol pieces (names follow csrc/cpppd_hot_kernels.cuh)
    def comm_wait(self, kind_in, deadline):
        want = self.wait_stamp[kind_in] + (1 if kind_in == 0 else 0)
        for t in self.recv_peers[kind_in]:
            while self.flags[kind_in * self.world + t] < want:
                if time.time() > deadline:
                    raise TimeoutError("rank %d waits for stamp %d of kind %d from rank %d (has %d)" % (
                        self.rank, want, kind_in, t, self.flags[kind_in * self.world + t]))
                time.sleep(0)

    def comm_finish(self, kind_out):
        self.push_stamp[kind_out] += 1
        for t in self.send_peers[kind_out]:
            self.ranks[t].flags[kind_out * self.world + self.rank] = self.push_stamp[kind_out]
        if self.recv_peers[1 - kind_out]:
            self.wait_stamp[1 - kind_out] += 1

    def stall(self):
        if self.rng.random() < 0.15:
            time.sleep(self.rng.random() * 2e-3)

    def kernel(self, kind_out, deadline):
        st = self.st
        nrows = st.nloc if kind_out == 0 else st.mloc
        order = list(range((nrows + 31) // 32))
        self.rng.shuffle(order)
        for s in order:
            rows = slice(32 * s, min(32 * s + 32, nrows))
            if self.role[kind_out][s]:
                self.comm_wait(1 - kind_out, deadline)
            self.stall()
            if kind_out == 0:
                d = st.c[rows]
                if st.has_eq:
                    d = d + st.At_eq[rows] @ st.y
                if st.has_ineq:
                    d = d + st.At_in[rows] @ st.y
                x2 = st.x[rows] - st.T[rows] * d
                np.maximum(x2, st.lb[rows], x2)
                np.minimum(x2, st.ub[rows], x2)
                out = (1 + st.theta) * x2 - st.theta * st.x[rows]
                st.xbar[rows] = out
                st.x[rows] = x2
            else:
                r = st.A[rows] @ st.xbar - st.b[rows]
                out = st.y[rows] + st.sigma[rows] * r
                lo = max(st.m_eq_loc - rows.start, 0)
                out[lo:] = np.maximum(out[lo:], 0)
                st.y[rows] = out
            if self.role[kind_out][s] & 1:
                for i in range(rows.start, rows.stop):
                    for t, slot in self.send[kind_out].get(i, ()):
                        peer = self.ranks[t].st
                        (peer.xbar if kind_out == 0 else peer.y)[slot] = out[i - rows.start]
        self.comm_finish(kind_out)


def run(args, world, iters, granule=32, seed=0, timeout_s=120):
    """Returns assembled (x, y) after `iters` iterations of the fused protocol on `world` thread-ranks."""
    ranks = [FusedRank(args, r, world, granule, random.Random(seed * 100 + r)) for r in range(world)]
    st0 = ranks[0].st
    import scipy.sparse as sp
    from oracle.cpppd_oracle import one_sided_system

    c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub = args
    blocks = [sp.csr_matrix(a) for a in (a_eq if a_eq is not None and a_eq.shape[0] else None,
                                         one_sided_system(a_ineq, b_lower, b_upper)[0]) if a is not None]
    indptr = np.concatenate(([0], np.cumsum(np.concatenate([np.diff(b_.indptr) for b_ in blocks]))))
    indices = np.concatenate([b_.indices for b_ in blocks])
    gcols, grows = {}, {}
    for t in range(world):
        gcols[t], grows[t] = po.ghosts(indptr, indices, st0.part, t)
    for r in ranks:
        r.link(ranks, gcols, grows)
    deadline = time.time() + timeout_s
    errors = []

    def body(rk):
        try:
            for _ in range(iters):
                rk.kernel(0, deadline)
                rk.kernel(1, deadline)
        except Exception as e:  # surfaced by the caller
            errors.append(e)

    threads = [threading.Thread(target=body, args=(rk,)) for rk in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    x = np.zeros(st0.n)
    y = np.zeros(st0.m)
    for rk in ranks:
        x[rk.st.own_cols] = rk.st.x
        y[rk.st.own_rows] = rk.st.y[: rk.st.mloc]
    return x, y
