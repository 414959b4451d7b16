"""TEST INFRASTRUCTURE — CPU restatement (numpy/scipy) of the reference CP-PPD solver.

This file is the *oracle*: an independent, executable statement of what
``pysparselp/ChambollePockPPD.py:36-346`` computes, used only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs.  The product package (``pysparselp_b200``) never imports it.

Parity status: **pinned**.  ``tests/test_oracle_golden.py`` checks this restatement
against (a) the reference's own golden curves for the path
(``tests/netlib_curves_SC105.json``, ``tests/test_pott_segmentation_curves.json``,
``tests/test_l1_svm_results.json`` — copied values live in ``tests/golden/``) and
(b) traces minted from the unmodified reference by ``oracle/make_golden.py``
(bit-for-bit: x, y, energies and violation curves).

The arithmetic is delegated to the same third-party routines the reference uses
(scipy ``csr_matvec`` / ``csc_matvec``, numpy ufuncs and ``dot``), so on one
machine this oracle and the reference agree to the last bit:

* ``Aᵀy``   — reference ``y * A`` (``:206,216``) is ``A.T @ y`` i.e. ``csc_matvec``:
  a sequential scatter ``d[j] += a_ij * y_i`` in row order starting from 0.0.
* ``A x``   — ``csr_matvec``: sequential sum per row in stored order from 0.0.
* ``c·x``   — ``numpy.dot`` (pairwise/BLAS order; only reproducible to rounding).
"""
import time

import numpy as np
import scipy.sparse as sp


def one_sided_system(a_ineq, b_lower, b_upper):
    """Inequality block as ``A x <= b``  (reference ``ChambollePockPPD.py:74-88``).

    Rows with a finite upper bound come first (original order), then the negated
    rows with a finite lower bound (original order).  When only one side exists the
    reference keeps/negates the *whole* matrix without filtering rows, and then
    trips its own shape assertion (``:141``) if some row had no finite bound; the
    same assertion is raised here.
    """
    if a_ineq is None or b_lower is None:
        return a_ineq, b_upper
    keep_up = np.flatnonzero(b_upper != np.inf)
    keep_lo = np.flatnonzero(b_lower != -np.inf)
    if keep_up.size and keep_lo.size:
        a = sp.vstack((a_ineq[keep_up, :], -a_ineq[keep_lo, :])).tocsr()
    elif keep_lo.size:
        a = -a_ineq
    else:
        a = a_ineq
    b = np.concatenate((b_upper[keep_up], -b_lower[keep_lo]))
    assert a.shape[0] == b.size, "rows without a finite bound: reference asserts here (:141)"
    return a, b


def _abs_pow(a, p):
    out = a.copy()
    out.data = np.abs(out.data) ** p
    return out


class CpPpdOracle:
    """State machine form of the solver so tests can step it and look inside."""

    def __init__(self, c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1):
        if a_eq is not None and a_eq.shape[0] == 0:  # :70-72
            a_eq, beq = None, None
        a_ineq, b_ineq = one_sided_system(a_ineq, b_lower, b_upper)
        self.c, self.lb, self.ub = c, lb, ub
        self.a_eq, self.b_eq, self.a_ineq, self.b_ineq = a_eq, beq, a_ineq, b_ineq
        self.alpha, self.theta = alpha, theta
        self.n = c.size
        assert lb.size == self.n and ub.size == self.n  # :95-96
        self.x = x0.copy() if x0 is not None else np.zeros(self.n)  # :91-94
        self.xbar = self.x
        self.trivial = a_eq is None and a_ineq is None
        if self.trivial:
            return
        # primal preconditioner T  (:122-153): column sums of |A|^(2-alpha)
        col = 0
        if a_eq is not None:
            assert a_eq.shape[1] == self.n and a_eq.shape[0] == beq.size
            col = col + _abs_pow(a_eq, 2 - alpha).T @ np.ones(a_eq.shape[0])
        if a_ineq is not None:
            assert a_ineq.shape[1] == self.n and a_ineq.shape[0] == b_ineq.size
            col = col + _abs_pow(a_ineq, 2 - alpha).T @ np.ones(a_ineq.shape[0])
        col = np.asarray(col, dtype=np.float64)
        self.t_replaced = col == 0
        col[self.t_replaced] = 1
        self.diag_t = 1 / col
        # dual preconditioners Sigma (:158-179): row sums of |A|^alpha
        self.sig_eq = self.sig_ineq = None
        self.y_eq = self.y_ineq = None
        if a_eq is not None:
            rs = _abs_pow(a_eq, alpha) @ np.ones(self.n)
            rs[rs == 0] = 1
            self.sig_eq = 1 / rs
            self.y_eq = np.zeros(a_eq.shape[0])
            self.at_eq = a_eq.T  # csc view: (at_eq @ y) runs csc_matvec exactly like `y * a_eq`
        if a_ineq is not None:
            rs = _abs_pow(a_ineq, alpha) @ np.ones(self.n)
            rs[rs == 0] = 1
            self.sig_ineq = 1 / rs
            self.y_ineq = np.zeros(a_ineq.shape[0])
            self.at_ineq = a_ineq.T
        self.best_energy = np.inf
        self.best_integer = None
        self.niter = 0

    # -- one iteration, split where the reference interleaves the stats block ------
    def primal_step(self):
        """:198-240 — returns nothing; leaves d, x, xbar, r_eq, r_ineq on self."""
        d = self.c
        if self.a_eq is not None:
            d = d + self.at_eq @ self.y_eq
        if self.a_ineq is not None:
            d = d + self.at_ineq @ self.y_ineq
        x2 = self.x - self.diag_t * d
        np.maximum(x2, self.lb, x2)
        np.minimum(x2, self.ub, x2)
        xbar_prev = self.xbar
        self.xbar = (1 + self.theta) * x2 - self.theta * self.x
        self.diff_xbar = xbar_prev - self.xbar
        self.x = x2
        self.d = d
        if self.a_eq is not None:
            self.r_eq = self.a_eq @ self.xbar - self.b_eq
        if self.a_ineq is not None:
            self.r_ineq = self.a_ineq @ self.xbar - self.b_ineq

    def stats(self, force_integer=False):
        """:248-291 — energies, violations, best-integer bookkeeping."""
        x, c = self.x, self.c
        e1 = c.dot(x)
        x4 = self.lb.copy()
        neg = self.d < 0
        x4[neg] = self.ub[neg]
        e2 = c.dot(x4)
        mv_eq = 0
        if self.a_eq is not None:
            e1 += self.y_eq.dot(self.a_eq @ x - self.b_eq)
            e2 += self.y_eq.dot(self.a_eq @ x4 - self.b_eq)
            mv_eq = np.max(np.abs(self.r_eq))
        if self.a_ineq is not None:
            e1 += self.y_ineq.dot(self.a_ineq @ x - self.b_ineq)
            e2 += self.y_ineq.dot(self.a_ineq @ x4 - self.b_ineq)
        xr = np.round(x) if force_integer else x
        e_r = c.dot(xr)
        mv_eq_r = np.max(np.abs(self.a_eq @ xr - self.b_eq)) if self.a_eq is not None else 0
        # the reference overwrites max_violated_inequality with the value at x_rounded (:283)
        mv_ineq = np.max(self.a_ineq @ xr - self.b_ineq)
        if mv_eq_r == 0 and mv_ineq <= 0 and e_r < self.best_energy:
            self.best_energy = e_r
            self.best_integer = xr
        return dict(
            niter=self.niter, energy1=e1, energy2=e2, max_violated_equality=mv_eq,
            max_violated_inequality=mv_ineq, energy_rounded=e_r, max_violated_equality_rounded=mv_eq_r,
            frac_zero_xbar=float(np.mean(self.xbar == 0)), frac_zero_diff=float(np.mean(self.diff_xbar == 0)),
        )

    def dual_step(self):
        """:333-343."""
        if self.a_eq is not None:
            self.y_eq = self.y_eq + self.sig_eq * self.r_eq
        if self.a_ineq is not None:
            self.y_ineq = self.y_ineq + self.sig_ineq * self.r_ineq
            np.maximum(self.y_ineq, 0, self.y_ineq)
        self.niter += 1


def chambolle_pock_ppd_oracle(
    c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1, nb_max_iter=100,
    callback_func=None, max_time=None, save_problem=False, force_integer=False, nb_iter_plot=10,
    trace=None, state_out=None, stats_enabled=True,
):
    """Same signature/return as the reference function (``:36-54,:344-346``).

    ``trace`` (list) receives one dict per stats block; ``state_out`` (dict) receives
    the final ``x, xbar, y_eq, y_ineq, diag_t, sig_eq, sig_ineq, niter``.
    ``stats_enabled=False`` skips the stats arithmetic (timing the bare loop).
    """
    start = time.perf_counter()
    o = CpPpdOracle(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0, alpha, theta)
    if o.trivial:  # :147-151 — note: bare vector, not a tuple
        x = np.zeros_like(lb)
        x[c > 0] = lb[c > 0]
        x[c < 0] = ub[c < 0]
        return x
    while o.niter < nb_max_iter:
        o.primal_step()
        if stats_enabled and o.niter % nb_iter_plot == 0:
            elapsed = time.perf_counter() - start
            if max_time is not None and elapsed > max_time:
                break
            s = o.stats(force_integer)
            if trace is not None:
                trace.append(s)
            if callback_func is not None:
                callback_func(o.niter, o.x, s["energy1"], s["energy2"], elapsed,
                              s["max_violated_equality"], s["max_violated_inequality"])
        o.dual_step()
    if state_out is not None:
        state_out.update(x=o.x, xbar=o.xbar, y_eq=o.y_eq, y_ineq=o.y_ineq, diag_t=o.diag_t,
                         sig_eq=o.sig_eq, sig_ineq=o.sig_ineq, niter=o.niter,
                         t_replaced=o.t_replaced)
    best = o.best_integer[: o.n] if o.best_integer is not None else None
    return o.x[: o.n], best
