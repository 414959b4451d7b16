"""Multi-GPU parity worker: `torchrun --nproc-per-node N tests/dist_worker.py` (one rank per GPU).

Every rank solves the same LPs through the public API in distributed mode and checks
  * iterates (x, y, T, Sigma) bit-identical to the goldens minted from the reference
    (owner-computes + ghost exchange never splits a row or column sum across GPUs),
  * stats curves within 1e-6 relative,
  * the partition (owned / ghost ids of this rank) equal to oracle/partition_oracle.py.
Prints DIST_WORKER_OK on rank 0 when everything passed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist

    from conftest import CASE_PARAMS, GOLDEN_CASES, case_args
    from oracle import partition_oracle as po
    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd, one_sided_rows, stack_operator
    from test_gpu_parity import assert_curves_close

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()

    for name in GOLDEN_CASES:
        args, g = case_args(name)
        kw = CASE_PARAMS.get(name, {})
        trace = []
        x, best, solver = chambolle_pock_ppd(
            *args, nb_max_iter=100, nb_iter_plot=10, return_solver=True, partition_granule=32,
            callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)), **kw)
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        info = solver.info()
        assert info["world_size"] == world and info["rank"] == rank
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]), name
            assert np.array_equal(y, y_gold), name
            assert np.array_equal(T, g["diag_t"]), name
        else:
            assert np.allclose(x, g["x_100"], rtol=1e-9, atol=0) and np.allclose(y, y_gold, rtol=1e-9, atol=1e-300)
        assert_curves_close(np.array(trace), g["trace_10"])
        # partition of this rank against the numpy restatement
        c, a_eq, beq, a_in, b_lo, b_up, lb, ub = args
        a_in1, b_in1 = one_sided_rows(a_in, b_lo, b_up)
        A, b, m_eq = stack_operator(a_eq if a_eq is not None and a_eq.shape[0] else None, beq, a_in1, b_in1, c.size)
        part = po.partition(A.indptr, A.indices, A.shape[1], m_eq, world, granule=32)
        own_c, ghost_c = solver.layout(columns=True)
        own_r, ghost_r = solver.layout(columns=False)
        assert np.array_equal(own_c, part["col_order"][part["col_start"][rank]: part["col_start"][rank + 1]]), name
        assert np.array_equal(own_r, part["row_order"][part["row_start"][rank]: part["row_start"][rank + 1]]), name
        gc, gr = po.ghosts(A.indptr, A.indices, part, rank)
        assert np.array_equal(ghost_c, gc) and np.array_equal(ghost_r, gr), name
        solver.close()
        # force_integer bookkeeping (best integer solution assembled across ranks)
        x, best = chambolle_pock_ppd(*args, nb_max_iter=300, nb_iter_plot=20, force_integer=True, **kw)
        if g["best_300_fi"].size:
            assert best is not None and np.array_equal(best, g["best_300_fi"]), name
        else:
            assert best is None
        if rank == 0:
            print("case %-20s ok on %d GPUs: owned cols %d ghost cols %d ghost rows %d" % (
                name, world, own_c.size, ghost_c.size, ghost_r.size), flush=True)

    # mid-size Potts: halo must be thin, iterates equal to the single-GPU oracle
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    lp = generators.potts_lp(256)
    args = generators.lp_args(lp)
    xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=30, nb_iter_plot=1000)
    x, _, solver = chambolle_pock_ppd(*args, nb_max_iter=30, nb_iter_plot=1000, return_solver=True)
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo)
    assert info["n_ghost"] <= 2 * 256 + 64 and info["m_ghost"] <= 8 * 256 + 64, info
    # the other halo transport (NCCL send/recv instead of the peer-memory push / wait kernels), the compressed
    # storage and every kernel variant must give the same bits
    from pysparselp_b200 import _cabi

    for flags, variant in ((_cabi.FLAG_NO_P2P, 0), (_cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS, 0), (0, 3), (0, 5)):
        x2, _, solver = chambolle_pock_ppd(*args, nb_max_iter=30, nb_iter_plot=1000, return_solver=True, flags=flags,
                                           kernel_variant=variant)
        solver.close()
        assert np.array_equal(x2, xo), (flags, variant)
    # x0 warm start in distributed mode
    rng = np.random.default_rng(3)
    args, _ = case_args("random_small")
    x0 = rng.standard_normal(args[0].size)
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    x, _ = chambolle_pock_ppd(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    assert np.array_equal(x, xo)
    # banded operands across ranks (csrc/cpppd_banded.cuh): a random LP without locality takes the balanced split in
    # original order; x, y after 6 iterations must hash to the digest the C port minted (tests/golden/bench_digests.json)
    import json

    import bench

    os.environ["CPPPD_BAND_WINDOW_MB"] = "0.25"
    table = json.load(open(bench.DIGESTS))
    for size in (200000, 2000000):
        lp, _ = bench.build_workload("random", size, pinned=False)
        largs = generators.lp_args(lp)
        for flags in (_cabi.FLAG_BANDED, 0, _cabi.FLAG_BANDED | _cabi.FLAG_NO_P2P):
            x, _, solver = chambolle_pock_ppd(*largs, nb_max_iter=bench.DIGEST_ITERS, nb_iter_plot=1000, return_solver=True,
                                              flags=flags)
            y = solver.get_y()
            info = solver.info()
            solver.close()
            assert bench.iterate_digest(x, y) == table[bench.workload_name("random", size)]["sha256"], (size, flags)
            if flags & _cabi.FLAG_BANDED:
                assert info["balanced_split"] == 1 and info["band_in_use"] == [1, 1], info
        if rank == 0:
            print("random LP %d x %d banded on %d GPUs: digest ok (windows %r)" % (size, 2 * size, world, info["band_windows"]),
                  flush=True)
    os.environ.pop("CPPPD_BAND_WINDOW_MB")
    dist.barrier()
    if rank == 0:
        print("DIST_WORKER_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
