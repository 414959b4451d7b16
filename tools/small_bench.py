"""Scratch: it/s of the two latency-bound configs (Potts 50x50, SC105) on the graph path and the persistent kernels."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import bench
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import make_solver

a = argparse.Namespace(small_iters=int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
for _ in range(2):
    print(json.dumps(bench.small_configs(a, generators, make_solver)), flush=True)
for size in (24, 50, 64, 72, 96):
    out = {"potts": size}
    for label, flags in (("cuda_graphs", 4096), ("persistent", 0), ("persistent_reorder", 8)):
        s = make_solver(*generators.lp_args(generators.potts_lp(size)), flags=flags)
        s.iterate(2000); s.sync()
        out[label] = round(5000 / (s.time_iterations(5000) * 1e-3))
        out[label + "_kind"] = s.info()["tiny_persistent"]
        s.close()
    print(json.dumps(out), flush=True)
