import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import make_solver
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lp = generators.potts_lp(size)
s = make_solver(*generators.lp_args(lp), flags=flags)
s.iterate(20); s.sync()
info = s.info()
p, d = [v / 32 for v in s.time_kernels(32)]
ms = s.time_iterations(100) / 100
print("size %d flags %d: primal %.4f ms dual %.4f ms | iteration %.4f ms = %.1f it/s, algorithmic %.0f GB/s, actual %.0f GB/s" % (
    size, flags, p, d, ms, 1e3 / ms, info["bytes_per_iteration_algorithmic"] / ms / 1e6, info["bytes_per_iteration_actual"] / ms / 1e6))
