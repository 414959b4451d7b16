"""TEST INFRASTRUCTURE — dry-run the `-m gpu` tests on a machine without a GPU:

    CPPPD_EMULATE_GPU_TESTS=1 python -m pytest tests -m gpu -p emul.patch_plugin -q

The plugin swaps, inside the test process only, the product's `CpPpdSolver` for an adapter over the
CPU-emulated library (tests/emul/cabi_driver.py), so the GPU test code itself (and the Python front it calls)
is exercised before a GPU run is spent on it.  It proves nothing about the hardware path; the real
`-m gpu` run on a B200 does not load this plugin.
"""
import pysparselp_b200.ChambollePockPPD as front

from .cabi_driver import EmulatedSolver


class _Adapter(EmulatedSolver):
    def __init__(self, c, a, m_eq, b, lb, ub, x0=None, alpha=1, theta=1, device=None, flags=0, process_group=None,
                 partition_granule=0, kernel_variant=0, long_row_threshold=0):
        super().__init__(c, a, m_eq, b, lb, ub, x0=x0, alpha=alpha, theta=theta, flags=flags,
                         partition_granule=partition_granule, kernel_variant=kernel_variant,
                         long_row_threshold=long_row_threshold)


def pytest_configure(config):
    front.CpPpdSolver = _Adapter
