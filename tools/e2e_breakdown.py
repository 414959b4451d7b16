"""Where does an end-to-end call spend its time? (scratch tool)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_workload
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import make_solver, chambolle_pock_ppd

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lp, keep = build_workload("potts", size, pinned=True)
args = generators.lp_args(lp)
for flags in (0, 8, 0, 8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = make_solver(*args, flags=flags)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s.iterate(50); s.sync(); t2 = time.perf_counter()
    ms = s.time_iterations(200) / 200
    t3 = time.perf_counter()
    x = s.get_x(); t4 = time.perf_counter()
    s.primal_step(True); s.stats_step(False); st = s.read_stats(); s.dual_step(); s.sync(); t5 = time.perf_counter()
    s.close(); t6 = time.perf_counter()
    print("flags %d: make_solver %.3f s | first 50 its (graph build) %.3f s | %.4f ms/it | get_x %.3f s | stats iteration %.4f s | close %.3f s" % (
        flags, t1 - t0, t2 - t1, ms, t4 - t3, t5 - t4, t6 - t5), flush=True)
for flags in (0, 8):
    for _ in range(2):
        t0 = time.perf_counter()
        x, best = chambolle_pock_ppd(*args, nb_max_iter=500, nb_iter_plot=500, flags=flags)
        t1 = time.perf_counter()
        print("flags %d: e2e 500 iterations %.3f s -> %.1f it/s" % (flags, t1 - t0, 500 / (t1 - t0)), flush=True)
