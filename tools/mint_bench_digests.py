"""Mint tests/golden/bench_digests.json: sha256 of (x, y) after DIGEST_ITERS iterations of the plain-C oracle port
(oracle/cpppd_oracle.c — itself pinned bit for bit to the goldens minted from the unmodified reference,
tests/test_oracle_golden.py) on the bench workloads.  bench.py compares the digest of its timed configuration (any
number of GPUs, any halo transport, any storage) with these: `parity: ok` in the JSON line, non-zero exit otherwise.

    python tools/mint_bench_digests.py [potts:4096 random:20000000 potts:1024 random:2000000 ...]

Host memory: about 15 GB for the two full-size workloads; a few minutes on 8 cores.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import c_port  # noqa: E402
from oracle.c_port import COracle  # noqa: E402
from pysparselp_b200 import generators  # noqa: E402


def main():
    jobs = sys.argv[1:] or ["potts:4096", "random:20000000", "potts:1024", "potts:256", "random:2000000", "random:200000",
                            "potts:24", "random:2000"]  # (the last two: the CPU dry run of bench.py)
    path = bench.DIGESTS
    table = json.load(open(path)) if os.path.isfile(path) else {}
    threads = c_port.set_threads()
    for job in jobs:
        kind, size = job.split(":")
        size = int(size)
        name = bench.workload_name(kind, size)
        t0 = time.time()
        lp, _ = bench.build_workload(kind, size, pinned=False)
        co = COracle(*generators.lp_args(lp))
        co.iterate(bench.DIGEST_ITERS)
        table[name] = {"sha256": bench.iterate_digest(co.x, co.y), "iterations": bench.DIGEST_ITERS, "n": int(co.n), "m": int(co.m),
                       "minted_by": "tools/mint_bench_digests.py: oracle/cpppd_oracle.c (OpenMP, %d threads)" % threads}
        if kind == "l1svm":  # long weight columns: compared within 1e-9, not bit for bit
            table[name]["fingerprint"] = bench.iterate_fingerprint(co.x, co.y)
        print(name, table[name]["sha256"], "%.1f s" % (time.time() - t0), flush=True)
        del co, lp
        with open(path, "w") as f:
            json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
