import sys; sys.path.insert(0,'.')
import numpy as np
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import make_solver
from oracle.cpppd_oracle import CpPpdOracle
size = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lp = generators.potts_lp(size)
args = generators.lp_args(lp)
s = make_solver(*args)
print(s.info())
o = CpPpdOracle(*args)
T, sig = s.get_preconditioners()
print("T equal", np.array_equal(T, o.diag_t), "sigma equal", np.array_equal(sig, o.sig_ineq))
s.iterate(1)
o.primal_step(); o.dual_step()
print("x equal", np.array_equal(s.get_x(), o.x), "y equal", np.array_equal(s.get_y(), o.y_ineq))
s.primal_step(keep_d=True)
s.sync()
o.primal_step()
print("x equal after keep_d primal", np.array_equal(s.get_x(), o.x), np.array_equal(s.get_d(), o.d))
