"""CPU: the stamp / flag protocol of the halo exchange fused into k_primal / k_dual, played by one
thread per rank with random slice order and random stalls (oracle/fused_protocol_sim.py)."""
import numpy as np
import pytest

from conftest import case_args
from oracle import fused_protocol_sim as sim


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["potts50", "sc105", "random_small"])
def test_fused_protocol_is_deadlock_free_and_exact(name, world):
    args, g = case_args(name)
    iters = 40
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    st = {}
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=iters, nb_iter_plot=10**6, state_out=st)
    yo = np.concatenate([v for v in (st["y_eq"], st["y_ineq"]) if v is not None])
    for seed in (0, 1):
        x, y = sim.run(args, world, iters, seed=seed)
        assert np.array_equal(x, xo) and np.array_equal(y, yo)
