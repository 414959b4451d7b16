"""CPU: the real CUDA sources of k_primal / k_dual (pysparselp_b200/csrc/cpppd_hot_kernels.cuh), compiled
for the host through tests/emul/cuda_shim.h, against the goldens minted from the reference.  Bit-exact,
like the GPU parity tests — this is the kernel logic, run without a GPU."""
import numpy as np
import pytest

from pysparselp_b200 import _cabi

from conftest import CASE_PARAMS, GOLDEN_CASES, case_args
from emul.harness import EmulSolver, lib


def test_emulator_uses_the_product_constants():
    assert lib().emul_constants(0) == 32 and lib().emul_constants(1) == 256
    assert lib().emul_constants(2) == 0x40000000 and lib().emul_constants(3) == -2**31


def test_emulator_sees_every_kernel_variant():
    from pysparselp_b200 import _cabi

    assert lib().emul_num_variants() == _cabi.KERNEL_VARIANTS


@pytest.mark.parametrize("variant", range(_cabi.KERNEL_VARIANTS))
@pytest.mark.parametrize("compressed", [False, True])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hot_kernels_bit_exact_on_cpu(name, compressed, variant):
    """every variant of the two kernels (loop flavour x registers) must give the same bits"""
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    s = EmulSolver(*args, value_dict=compressed, const_vectors=compressed, variant=variant, **kw)
    if compressed and name == "potts50":
        assert s.dict is not None and s.dict.size == 2  # the +-1 matrix really takes the dictionary path
    s.iterate(50)
    s.primal(write_d=True)  # the stats-iteration variant of the kernel (keeps d)
    s.dual()
    s.iterate(49)
    y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
    assert np.array_equal(s.x, g["x_100"])
    assert np.array_equal(s.y, y_gold)


def test_hot_kernels_with_x0_and_ragged_sizes():
    import scipy.sparse as sp

    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    rng = np.random.default_rng(5)
    n = 37
    a = sp.random(45, n, density=0.1, random_state=2, format="csr")
    a.data = np.round(a.data * 6 - 3, 1)
    a = a.tolil()
    a[3, :] = 0
    a[:, 5] = 0
    a = a.tocsr()
    c = rng.standard_normal(n)
    lb, ub = -np.ones(n), np.ones(n)
    b_up = rng.random(45)
    x0 = rng.standard_normal(n)
    args = (c, sp.csr_matrix((0, n)), np.empty(0), a, None, b_up, lb, ub)
    st = {}
    xo, _ = chambolle_pock_ppd_oracle(*args, x0=x0, nb_max_iter=60, nb_iter_plot=10**6, state_out=st)
    for variant in range(_cabi.KERNEL_VARIANTS):
        s = EmulSolver(*args, x0=x0, variant=variant)
        s.iterate(60)
        assert np.array_equal(s.x, xo) and np.array_equal(s.y, st["y_ineq"]) and np.array_equal(s.xbar, st["xbar"])
