"""CPU: the multi-GPU code of the real library (partition, ghost lists, halo exchange over "peer memory",
NCCL transport, distributed stats, vector assembly) with world_size 2 and 3 — every rank is an OS thread
running the CPU-emulated library (tests/emul), peer memory is the shared address space, NCCL is
tests/emul/emul_nccl.cpp.  Same assertions as tests/dist_worker.py makes on real GPUs."""
import numpy as np
import pytest

from conftest import CASE_PARAMS, case_args
from emul.cabi_driver import make_emulated_solver, run_ranks
from oracle import partition_oracle as po
from pysparselp_b200 import _cabi
from pysparselp_b200.ChambollePockPPD import one_sided_rows, run_schedule, stack_operator

TRANSPORTS = {"push_wait_kernels": 0, "fused_in_kernels": _cabi.FLAG_FUSED_HALO, "nccl": _cabi.FLAG_NO_P2P}


def solve_on_ranks(args, world, flags, nb_max_iter, nb_iter_plot, force_integer=False, **kw):
    def body(rank, world, comm_id):
        trace = []
        solver = make_emulated_solver(*args, flags=flags, partition_granule=32, rank=rank, world=world,
                                      comm_id=comm_id, **kw)
        x, best = run_schedule(solver, nb_max_iter, lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)),
                               None, force_integer, nb_iter_plot)
        out = dict(x=x, best=best, trace=np.array(trace), y=solver.get_y(), T=solver.get_preconditioners()[0],
                   cols=solver.layout(True), rows=solver.layout(False), info=solver.info())
        solver.close()
        return out

    return run_ranks(world, body)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("transport", list(TRANSPORTS))
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["potts50", "sc105", "random_small", "afiro", "kb2"])
def test_multi_rank_iterates_bit_identical(name, world, transport):
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    res = solve_on_ranks(args, world, TRANSPORTS[transport], 100, 10, **kw)
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    c, a_eq, beq, a_in, b_lo, b_up, lb, ub = args
    a_in1, b_in1 = one_sided_rows(a_in, b_lo, b_up)
    A, b, m_eq = stack_operator(a_eq if a_eq is not None and a_eq.shape[0] else None, beq, a_in1, b_in1, c.size)
    part = po.partition(A.indptr, A.indices, A.shape[1], m_eq, world, granule=32)
    for rank, r in enumerate(res):
        assert np.array_equal(r["x"], g["x_100"]) and np.array_equal(r["y"], y_gold) and np.array_equal(r["T"], g["diag_t"])
        assert np.allclose(r["trace"], g["trace_10"], rtol=1e-6, atol=1e-9, equal_nan=True)
        assert r["info"]["world_size"] == world and r["info"]["rank"] == rank
        own_c, ghost_c = r["cols"]
        own_r, ghost_r = r["rows"]
        assert np.array_equal(own_c, part["col_order"][part["col_start"][rank]: part["col_start"][rank + 1]])
        assert np.array_equal(own_r, part["row_order"][part["row_start"][rank]: part["row_start"][rank + 1]])
        gc, gr = po.ghosts(A.indptr, A.indices, part, rank)
        assert np.array_equal(ghost_c, gc) and np.array_equal(ghost_r, gr)


@pytest.mark.timeout(900)
def test_multi_rank_force_integer_and_compression():
    args, g = case_args("potts50")
    flags = _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS
    res = solve_on_ranks(args, 2, flags, 300, 20, force_integer=True)
    for r in res:
        assert np.array_equal(r["x"], g["x_300_fi"])
        assert r["best"] is not None and np.array_equal(r["best"], g["best_300_fi"])
        assert r["info"]["value_bytes"] == 0


@pytest.mark.timeout(900)
def test_multi_rank_device_curves():
    """cpppd_set_ground_truth with world_size 2: every rank sums the entries it owns, the stats all-gather
    adds them up — distances and the bound violation equal the numpy evaluation on the assembled x."""
    args, g = case_args("potts50")
    gt = g["ground_truth"].astype(float).ravel()
    idx = np.arange(gt.size)
    lb, ub = args[6], args[7]

    def body(rank, world, comm_id):
        solver = make_emulated_solver(*args, partition_granule=32, rank=rank, world=world, comm_id=comm_id)
        solver.set_ground_truth(idx, gt)
        seen = []

        def on_stats(niter, st, elapsed):
            x = solver.get_x()  # collective: every rank calls it at the same point
            seen.append((st["distance_to_ground_truth"], float(np.mean(np.abs(gt - x[idx]))),
                         st["distance_to_ground_truth_rounded"], float(np.mean(np.abs(gt - np.round(x[idx])))),
                         st["max_bound_violation"], float(max(np.max(lb - x), np.max(x - ub)))))

        run_schedule(solver, 60, None, None, False, 20, stats_func=on_stats)
        solver.close()
        return np.array(seen)

    for seen in run_ranks(2, body):
        assert seen.shape == (3, 6)
        assert np.allclose(seen[:, 0], seen[:, 1], rtol=1e-13, atol=1e-16)
        assert np.allclose(seen[:, 2], seen[:, 3], rtol=1e-13, atol=1e-16)
        assert np.array_equal(seen[:, 4], seen[:, 5])


@pytest.mark.timeout(900)
def test_multi_rank_autotune_and_forced_variants(monkeypatch):
    """Kernel variants are timed before the peer-memory rendezvous; ranks may end up with different variants and
    still produce the same bits."""
    monkeypatch.setenv("CPPPD_AUTOTUNE_MIN_NNZ", "0")
    args, g = case_args("potts50")
    for r in solve_on_ranks(args, 2, 0, 100, 10):
        assert r["info"]["autotuned"] == 1
        assert np.array_equal(r["x"], g["x_100"])
    for r in solve_on_ranks(args, 3, 0, 100, 10, kernel_variant=4):
        assert r["info"]["autotuned"] == 0 and r["info"]["dual_variant"] == 4
        assert np.array_equal(r["x"], g["x_100"])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("transport", ["push_wait_kernels", "fused_in_kernels"])
def test_halo_wait_gives_up_instead_of_hanging(transport, monkeypatch):
    """A neighbour that never delivers its halo must surface as CPPPD_ERR_COMM after CPPPD_HALO_TIMEOUT_S, not hang
    the device."""
    import time

    monkeypatch.setenv("CPPPD_HALO_TIMEOUT_S", "0.5")
    args, _ = case_args("potts50")

    def body(rank, world, comm_id):
        solver = make_emulated_solver(*args, flags=TRANSPORTS[transport], partition_granule=32, rank=rank, world=world,
                                      comm_id=comm_id)
        if rank == 1:  # this rank never iterates
            time.sleep(3.0)
            solver.handle = None  # (its destroy would wait for rank 0, which has given up)
            return "idle"
        t0 = time.time()
        try:
            solver.iterate(3)
            solver.sync()
        except _cabi.CpppdError as e:
            assert e.code == -5 and "timed out" in str(e)
            return time.time() - t0
        finally:
            solver.close()
        return None

    res = run_ranks(2, body)
    assert res[1] == "idle" and res[0] is not None and res[0] < 30.0


@pytest.mark.timeout(900)
@pytest.mark.parametrize("transport", ["push_wait_kernels", "nccl", "fused_in_kernels"])
def test_multi_rank_long_rows(transport):
    """Long rows / columns with world_size 2: the long column sums gather ghost y entries, so they must run after
    the halo arrived (the fused transport is switched off for such LPs)."""
    args, g = case_args("l1svm")
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    res = solve_on_ranks(args, 2, TRANSPORTS[transport], 100, 10, long_row_threshold=64)
    assert sum(r["info"]["long_cols"] for r in res) == 9  # the weight columns, wherever they live
    for r in res:
        assert np.max(np.abs(r["x"] - g["x_100"])) <= 1e-9 * np.max(np.abs(g["x_100"]))
        assert np.max(np.abs(r["y"] - y_gold)) <= 1e-9 * np.max(np.abs(y_gold))
        assert np.allclose(r["trace"], g["trace_10"], rtol=1e-6, atol=1e-9, equal_nan=True)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_balanced_split_for_patterns_without_locality(world):
    """Every row of the L1-SVM LP starts at a weight column: the locality buckets would give one rank everything.
    The partition then deals rows and columns out by prefix sums (oracle/partition_oracle.py); iterates stay
    bit-identical because whole rows and whole columns still live on one rank."""
    args, g = case_args("l1svm")
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    c, a_eq, beq, a_in, b_lo, b_up, lb, ub = args
    a_in1, b_in1 = one_sided_rows(a_in, b_lo, b_up)
    A, b, m_eq = stack_operator(None, None, a_in1, b_in1, c.size)
    part = po.partition(A.indptr, A.indices, A.shape[1], m_eq, world, granule=32)
    assert part["balanced_split"]
    res = solve_on_ranks(args, world, 0, 100, 10)
    for rank, r in enumerate(res):
        assert r["info"]["balanced_split"] == 1
        assert np.array_equal(r["x"], g["x_100"]) and np.array_equal(r["y"], y_gold)
        own_c, ghost_c = r["cols"]
        own_r, ghost_r = r["rows"]
        assert np.array_equal(own_c, part["col_order"][part["col_start"][rank]: part["col_start"][rank + 1]])
        assert np.array_equal(own_r, part["row_order"][part["row_start"][rank]: part["row_start"][rank + 1]])
        gc, gr = po.ghosts(A.indptr, A.indices, part, rank)
        assert np.array_equal(ghost_c, gc) and np.array_equal(ghost_r, gr)
        # nobody holds more than 1.5x its share of the entries, rows and columns alike
        assert r["info"]["nnz_local_rows"] * world * 2 <= 3 * A.nnz and r["info"]["nnz_local_cols"] * world * 2 <= 3 * A.nnz


@pytest.mark.timeout(900)
@pytest.mark.parametrize("transport", ["push_wait_kernels", "nccl"])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", ["random_small", "sc105", "kb2"])
def test_multi_rank_banded_operands(name, world, transport, monkeypatch):
    """Banded operands (csrc/cpppd_banded.cuh) on several ranks: patterns without locality take the balanced split, rows
    and columns keep their original order inside a rank, the windows are ranges of ORIGINAL ids (looked up through
    the local -> original maps), and the halo exchange follows the last window of every half-iteration.  Same bits
    as one GPU; the layout is the Python restatement's (keep_order)."""
    monkeypatch.setenv("CPPPD_BAND_WINDOW", "9")
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    res = solve_on_ranks(args, world, TRANSPORTS[transport] | _cabi.FLAG_BANDED, 100, 10, **kw)
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    c, a_eq, beq, a_in, b_lo, b_up, lb, ub = args
    a_in1, b_in1 = one_sided_rows(a_in, b_lo, b_up)
    A, b, m_eq = stack_operator(a_eq if a_eq is not None and a_eq.shape[0] else None, beq, a_in1, b_in1, c.size)
    part = po.partition(A.indptr, A.indices, A.shape[1], m_eq, world, granule=32, keep_order=True)
    banded = 0
    for rank, r in enumerate(res):
        assert np.array_equal(r["x"], g["x_100"]) and np.array_equal(r["y"], y_gold) and np.array_equal(r["T"], g["diag_t"])
        assert np.allclose(r["trace"], g["trace_10"], rtol=1e-6, atol=1e-9, equal_nan=True)
        own_c, ghost_c = r["cols"]
        own_r, ghost_r = r["rows"]
        assert np.array_equal(own_c, part["col_order"][part["col_start"][rank]: part["col_start"][rank + 1]])
        assert np.array_equal(own_r, part["row_order"][part["row_start"][rank]: part["row_start"][rank + 1]])
        if part["balanced_split"]:
            assert r["info"]["balanced_split"] == 1 and r["info"]["band_in_use"] == [1, 1]
            assert np.all(np.diff(own_c) > 0)  # original order inside the rank
            # dense halo: every foreign column / row is a ghost, peer by peer, each peer's slice in its own order
            assert r["info"]["dense_halo"] == 1
            others = [q for q in range(world) if q != rank]
            assert np.array_equal(ghost_c, np.concatenate(
                [part["col_order"][part["col_start"][q]: part["col_start"][q + 1]] for q in others]))
            assert np.array_equal(ghost_r, np.concatenate(
                [part["row_order"][part["row_start"][q]: part["row_start"][q + 1]] for q in others]))
            banded += 1
        else:  # locality buckets: the banded form is not built
            assert r["info"]["band_in_use"] == [0, 0]
    if name == "random_small":
        assert banded == world


@pytest.mark.timeout(300)
def test_distributed_vector_assembly_keeps_the_sign_of_zero():
    """cpppd_get_vector on several ranks assembles a vector by summing, over the ranks, a buffer that holds the owner's
    value and zero bits elsewhere.  The sum runs on the 64-bit patterns: a floating point sum would return +0.0 for an
    entry whose owner holds -0.0 (seen on hardware: the digest of a 2 M-variable random LP differed from the C port's
    while every element compared equal)."""
    args, g = case_args("random_small")
    n, m = args[0].size, g["y_eq"].size + g["y_ineq"].size
    xs = np.linspace(-1.0, 1.0, n)
    xs[::3] = -0.0
    ys = np.linspace(-2.0, 2.0, m)
    ys[1::4] = -0.0

    def body(rank, world, comm_id):
        solver = make_emulated_solver(*args, partition_granule=32, rank=rank, world=world, comm_id=comm_id)
        solver.set_x(xs)
        solver.set_y(ys)
        out = solver.get_x(), solver.get_y()
        solver.close()
        return out

    for x, y in run_ranks(3, body):
        assert np.array_equal(x, xs) and np.array_equal(np.signbit(x), np.signbit(xs))
        assert np.array_equal(y, ys) and np.array_equal(np.signbit(y), np.signbit(ys))


@pytest.mark.timeout(900)
def test_multi_rank_sparse_halo_of_banded_operands(monkeypatch):
    """CPPPD_FLAG_NO_DENSE_HALO: banded operands with the index-list halo (only the ghosts the pattern touches)."""
    monkeypatch.setenv("CPPPD_BAND_WINDOW", "9")
    args, g = case_args("random_small")
    res = solve_on_ranks(args, 3, _cabi.FLAG_BANDED | _cabi.FLAG_NO_DENSE_HALO, 100, 10)
    y_gold = np.concatenate([g["y_eq"], g["y_ineq"]])
    for r in res:
        assert r["info"]["dense_halo"] == 0 and r["info"]["band_in_use"] == [1, 1]
        assert np.array_equal(r["x"], g["x_100"]) and np.array_equal(r["y"], y_gold)


@pytest.mark.timeout(600)
def test_peer_memory_pool_of_a_kept_communicator(monkeypatch, capfd):
    """Solves made on one cpppd_comm (what the Python layer does) take xbar / y / stamps and the peer mappings from the
    communicator's pool: built by the first solve, reused by the second, rebuilt when a larger LP arrives, bypassed
    (private buffers) while another solver of the communicator is alive.  Same bits every time."""
    from emul.cabi_driver import emulated_comm, emulated_library

    monkeypatch.setenv("CPPPD_POOL_GRANULE", "256")  # (default 2 MB: every golden LP would fit the first pool)
    monkeypatch.setenv("CPPPD_POOL_TRACE", "1")
    small, g_small = case_args("sc105")
    large, g_large = case_args("potts50")

    def body(rank, world, comm_id):
        comm = emulated_comm(comm_id, rank, world)
        out = []
        try:
            for args in (small, small, large, small):
                solver = make_emulated_solver(*args, partition_granule=32, rank=rank, world=world, comm=comm)
                solver.iterate(100)
                out.append((solver.get_x(), solver.get_y()))
                solver.close()
            # two solvers alive at once: the second one must not touch the buffers of the first
            first = make_emulated_solver(*large, partition_granule=32, rank=rank, world=world, comm=comm)
            second = make_emulated_solver(*small, partition_granule=32, rank=rank, world=world, comm=comm)
            first.iterate(50)
            second.iterate(100)
            first.iterate(50)
            out.append((first.get_x(), first.get_y()))
            out.append((second.get_x(), second.get_y()))
            second.close()
            first.close()
        finally:
            emulated_library().cpppd_comm_destroy(comm)
        return out

    def gold_y(g):
        return np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])

    for world in (2, 3):
        for out in run_ranks(world, body):
            for (x, y), g in zip(out, (g_small, g_small, g_large, g_small, g_large, g_small)):
                assert np.array_equal(x, g["x_100"]) and np.array_equal(y, gold_y(g))
        trace = [l.split("] ")[1].split(" ")[0] for l in capfd.readouterr().err.splitlines() if "[cpppd pool rank 0]" in l]
        assert trace == ["built", "reused", "built", "reused", "reused", "busy:"], trace


@pytest.mark.timeout(300)
def test_time_out_is_one_decision_for_all_ranks():
    """run_schedule with max_time on two ranks whose clocks disagree (reference :243-247, the schedule of
    pysparselp_b200/ChambollePockPPD.py): rank 0's verdict is the verdict of every rank, so the ranks keep issuing the
    same collectives — a rank that broke out alone would call get_x (a collective) while the other one is inside the
    stats exchange, and both would hang."""
    import threading
    import time

    args, g = case_args("potts50")

    class Agreement:
        """agree() of the product's solver (one broadcast / all-reduce through torch.distributed) over two threads"""

        def __init__(self, world):
            self.barrier = threading.Barrier(world)
            self.votes = [False] * world

        def bind(self, solver, rank):
            def agree(flag, any_rank=False):
                self.votes[rank] = bool(flag)
                self.barrier.wait(timeout=60)
                verdict = any(self.votes) if any_rank else self.votes[0]
                self.barrier.wait(timeout=60)
                return verdict

            solver.agree, solver.world = agree, 2

    for late_rank, want_iterations in ((1, 40), (0, 0)):
        shared = Agreement(2)

        def body(rank, world, comm_id):
            solver = make_emulated_solver(*args, partition_granule=32, rank=rank, world=world, comm_id=comm_id)
            shared.bind(solver, rank)
            # the late rank believes the solve started 1000 s ago: on its own clock max_time = 500 s has passed
            start = time.perf_counter() - (1000.0 if rank == late_rank else 0.0)
            x, best = run_schedule(solver, 40, None, 500.0, False, 10, False, start)
            niter = solver.niter
            solver.close()
            return x, niter

        results = run_ranks(2, body)
        assert results[0][1] == results[1][1] == want_iterations
        assert np.array_equal(results[0][0], results[1][0])
