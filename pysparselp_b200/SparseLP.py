"""LP modeling layer + ``solve()`` dispatch for the B200 CP-PPD solver.

Host-side (CPU, numpy/scipy) mirror of the part of the reference's ``SparseLP`` class
that feeds the Chambolle-Pock path — same method names, argument meaning and
attributes, so code written against ``pysparselp.SparseLP`` runs unchanged:

* variables / constraints builders  (reference ``pysparselp/SparseLP.py:421-613``)
* ``remove_fixed_variables``         (``:632-674``)
* ``convert_to_one_sided_inequality_system`` (``:835-879``), ``convert_to_all_equalities`` (``:819-833``),
  ``convert_to_all_inequalities[_without_bounds]`` (``:881-928``)
* ``max_constraint_violation`` / ``check_solution`` (``:186-226``)
* ``solve(method='chambolle_pock_ppd', ...)`` (``:990-1002, :1243-1288, :1378-1383``)

Design differences (none observable through the API): constraint rows are kept as
a list of CSR blocks and concatenated once on first read, instead of growing three
numpy arrays with ``np.append`` per call (``:93-104``), which is quadratic for the
large generated LPs this build targets.

Only ``method='chambolle_pock_ppd'`` is implemented here: the other solvers of the
reference (ADMM, dual ascent, Mehrotra, external adapters) are outside the scope of
this build and raise ``NotImplementedError``.
"""
import copy
import time

import numpy as np
import scipy.sparse as sp

from .ChambollePockPPD import chambolle_pock_ppd, make_solver, run_schedule

solving_methods = ("chambolle_pock_ppd",)

# methods the reference knows (``SparseLP.py:45-56``) but this build does not provide
_reference_only_methods = (
    "osqp", "mehrotra", "scipy_simplex", "scipy_interior_point", "dual_coordinate_ascent",
    "dual_gradient_ascent", "admm", "admm2", "admm_blocks", "ECOS", "SCS", "CVXOPT",
)


def _is_plain_number(v):
    return isinstance(v, (int, float, np.floating, np.integer)) and not isinstance(v, bool)


def crd_matrix(cols, vals, broadcast=True):
    """CSR matrix with a constant number of candidate entries per row.

    ``m[i, cols[i, j]] = vals[i, j]``; entries whose value is exactly zero are not
    stored (reference ``SparseLP.py:127-159``).  Entry order inside a row is the
    order of ``cols`` (NOT sorted) — the solver kernels rely on that order for
    bit-reproducible row sums.
    """
    cols = np.asarray(cols)
    vals = np.asarray(vals)
    if cols.ndim != 2 or vals.ndim != 2:
        raise ValueError("cols and vals must be 2-D")
    srt = np.sort(cols, axis=1)
    dup_rows = np.flatnonzero(np.any(srt[:, 1:] == srt[:, :-1], axis=1))
    if dup_rows.size:
        raise ValueError(
            "the same variable appears twice in %d constraint(s): %s" % (dup_rows.size, dup_rows)
        )
    if broadcast:
        cols, vals = np.broadcast_arrays(cols, vals)
    if cols.shape != vals.shape:
        raise ValueError("cols and vals shapes differ")
    stored = vals != 0
    indptr = np.zeros(cols.shape[0] + 1, dtype=np.int64)
    np.cumsum(stored.sum(axis=1), out=indptr[1:])
    return sp.csr_matrix((vals[stored], cols[stored], indptr))


class _RowBlocks:
    """Append-only list of CSR row blocks with a cached concatenation."""

    def __init__(self):
        self.parts = []
        self.nrows = 0
        self._cat = None
        self._cat_cols = -1

    def append(self, block):
        block = sp.csr_matrix(block)
        if block.shape[0] == 0:
            return
        self.parts.append(block)
        self.nrows += block.shape[0]
        self._cat = None

    def replace(self, matrix):
        self.parts = []
        self.nrows = 0
        self._cat = None
        if matrix is not None:
            self.append(matrix)

    def as_csr(self, ncols):
        if self._cat is not None and self._cat_cols == ncols:
            return self._cat
        if not self.parts:
            out = sp.csr_matrix((0, ncols), dtype=np.float64)
        else:
            nnz = sum(p.nnz for p in self.parts)
            idx_dtype = np.int64 if (nnz >= 2**31 - 1 or ncols >= 2**31 - 1) else np.int32
            data = np.concatenate([p.data.astype(np.float64, copy=False) for p in self.parts])
            indices = np.concatenate([p.indices.astype(idx_dtype, copy=False) for p in self.parts])
            indptr = np.zeros(self.nrows + 1, dtype=idx_dtype)
            at, base = 1, 0
            for p in self.parts:
                indptr[at: at + p.shape[0]] = p.indptr[1:].astype(idx_dtype) + base
                at += p.shape[0]
                base += p.nnz
            width = max([ncols] + [p.shape[1] for p in self.parts])
            out = sp.csr_matrix((data, indices, indptr), shape=(self.nrows, width))
        self.parts = [out] if out.shape[0] else []
        self._cat, self._cat_cols = out, ncols
        return out


class SparseLP:
    """min c·x  s.t.  A_eq x = b_eq,  b_lower <= A_ineq x <= b_upper,  lb <= x <= ub."""

    def __init__(self):
        self.nb_variables = 0
        self.variables_dict = dict()
        self.upper_bounds = np.empty(0, dtype=np.float64)
        self.lower_bounds = np.empty(0, dtype=np.float64)
        self.costsvector = np.empty(0, dtype=np.float64)
        self.is_integer = np.empty(0, dtype=bool)
        self._ineq = _RowBlocks()
        self._eq = _RowBlocks()
        self.b_lower = np.empty(0, dtype=np.float64)
        self.b_upper = np.empty(0, dtype=np.float64)
        self.b_equalities = np.empty(0, dtype=np.float64)
        self.solver = "chambolle_pock"
        self.equalityConstraintNames = []
        self.inequalityConstraintNames = []
        self.solution = None

    # -- constraint matrices are materialised on demand --------------------------------
    @property
    def a_inequalities(self):
        return self._ineq.as_csr(self.nb_variables)

    @a_inequalities.setter
    def a_inequalities(self, m):
        self._ineq.replace(m)

    @property
    def a_equalities(self):
        return self._eq.as_csr(self.nb_variables)

    @a_equalities.setter
    def a_equalities(self, m):
        self._eq.replace(m)

    def nb_equality_constraints(self):
        return self._eq.nrows

    def nb_inequality_constraints(self):
        return self._ineq.nrows

    # -- constraint naming (reference :228-275) ------------------------------------------
    def start_constraint_name(self, name):
        if name:
            self._open_name = (name, self.nb_equality_constraints(), self.nb_inequality_constraints())

    def end_constraint_name(self, name):
        if not name:
            return
        opened, eq0, in0 = self._open_name
        assert opened == name
        if self.nb_equality_constraints() > eq0:
            self.equalityConstraintNames.append(
                {"name": name, "start": eq0, "end": self.nb_equality_constraints() - 1})
        if self.nb_inequality_constraints() > in0:
            self.inequalityConstraintNames.append(
                {"name": name, "start": in0, "end": self.nb_inequality_constraints() - 1})

    def find_inequality_constraints_from_name(self, name):
        return [d for d in self.inequalityConstraintNames if d["name"] == name]

    # -- feasibility helpers (reference :186-226) ------------------------------------------
    def get_variables_bounds(self):
        return None, self.lower_bounds, self.upper_bounds

    def max_constraint_violation(self, solution):
        worst = 0
        worst = max(worst, np.max(self.lower_bounds - solution))
        worst = max(worst, np.max(solution - self.upper_bounds))
        if self.nb_equality_constraints() > 0:
            worst = max(worst, np.max(np.abs(self.a_equalities * solution - self.b_equalities)))
        if self.nb_inequality_constraints() > 0:
            ax = self.a_inequalities * solution
            if self.b_upper is not None:
                worst = max(worst, np.max(ax - self.b_upper))
            if self.b_lower is not None:
                worst = max(worst, np.max(self.b_lower - ax))
        return worst

    def check_solution(self, solution, tol=1e-6):
        ok = True
        if self.lower_bounds is not None:
            ok = ok & (np.max(self.lower_bounds - solution) < tol)
        if self.upper_bounds is not None:
            ok = ok & (np.max(solution - self.upper_bounds) < tol)
        if self.nb_equality_constraints() > 0:
            ok = ok & (np.max(np.abs(self.a_equalities * solution - self.b_equalities)) < tol)
        if self.nb_inequality_constraints() > 0:
            ax = self.a_inequalities * solution
            if self.b_upper is not None:
                ok = ok & (np.max(ax - self.b_upper) < tol)
            if self.b_lower is not None:
                ok = ok & (np.max(self.b_lower - ax) < tol)
        return ok

    # -- variables (reference :421-509) ------------------------------------------------------
    def convert_bounds_to_vectors(self, shape, lower_bounds, upper_bounds):
        def expand(v, missing):
            if v is None:
                return np.full(shape, missing, dtype=np.float64)
            if _is_plain_number(v):
                return np.full(shape, v, dtype=np.float64)
            return v

        lower_bounds = expand(lower_bounds, -np.inf)
        upper_bounds = expand(upper_bounds, np.inf)
        assert tuple(np.shape(upper_bounds)) == tuple(shape)
        assert tuple(np.shape(lower_bounds)) == tuple(shape)
        return lower_bounds, upper_bounds

    def add_variables_array(self, shape, lower_bounds, upper_bounds, costs=0, name=None, is_integer=False):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        shape = tuple(int(s) for s in shape)
        count = int(np.prod(shape))
        indices = np.arange(count).reshape(shape) + self.nb_variables
        self.nb_variables += count
        if _is_plain_number(costs):
            costs = np.full(shape, costs, dtype=np.float64)
        assert tuple(costs.shape) == shape
        lower_bounds, upper_bounds = self.convert_bounds_to_vectors(shape, lower_bounds, upper_bounds)
        self.upper_bounds = np.concatenate((self.upper_bounds, np.ravel(upper_bounds).astype(np.float64)))
        self.lower_bounds = np.concatenate((self.lower_bounds, np.ravel(lower_bounds).astype(np.float64)))
        self.costsvector = np.concatenate((self.costsvector, np.ravel(costs).astype(np.float64)))
        self.is_integer = np.concatenate((self.is_integer, np.full(count, is_integer, dtype=bool)))
        if name:
            self.variables_dict[name] = indices
        return indices

    def set_bounds_on_variables(self, indices, lower_bounds, upper_bounds):
        flat = np.ravel(indices)
        self.lower_bounds[flat] = lower_bounds if _is_plain_number(lower_bounds) else np.ravel(lower_bounds)
        self.upper_bounds[flat] = upper_bounds if _is_plain_number(upper_bounds) else np.ravel(upper_bounds)

    def get_variables_indices(self, name):
        return self.variables_dict[name]

    def set_costs_variables(self, indices, costs):
        assert np.shape(costs) == np.shape(indices)
        self.costsvector[np.ravel(indices)] = np.ravel(costs)

    # -- constraints (reference :511-613) ------------------------------------------------------
    def add_equality_constraints_sparse(self, a, b):
        self._eq.append(sp.csr_matrix(a))
        self.b_equalities = np.concatenate((self.b_equalities, np.atleast_1d(b).astype(np.float64)))

    def add_inequality_constraints_sparse(self, a, lower_bounds=None, upper_bounds=None):
        """Append rows ``lower_bounds <= a x <= upper_bounds`` (scalar-equal bounds become equalities)."""
        a = sp.csr_matrix(a)
        rows = (a.shape[0],)
        scalar_equal = _is_plain_number(lower_bounds) and not isinstance(lower_bounds, np.floating) \
            and lower_bounds == upper_bounds
        lower_bounds, upper_bounds = self.convert_bounds_to_vectors(rows, lower_bounds, upper_bounds)
        if scalar_equal:
            self._eq.append(a)
            self.b_equalities = np.concatenate((self.b_equalities, lower_bounds))
            return
        self._ineq.append(a)
        if self.b_lower is None:  # after a one-sided conversion: keep one-sided bookkeeping consistent
            self.b_lower = np.full(self._ineq.nrows - a.shape[0], -np.inf)
        self.b_lower = np.concatenate((self.b_lower, np.asarray(lower_bounds, dtype=np.float64)))
        self.b_upper = np.concatenate((self.b_upper, np.asarray(upper_bounds, dtype=np.float64)))

    def add_equality_constraints(self, cols, vals, b):
        self.add_inequality_constraints(cols, vals, lower_bounds=b, upper_bounds=b)

    def add_soft_equality_constraints(self, cols, vals, b, coef_penalization):
        return self.add_soft_inequality_constraints(
            cols, vals, lower_bounds=b, upper_bounds=b, coef_penalization=coef_penalization)

    def add_inequality_constraints(self, cols, vals, lower_bounds=None, upper_bounds=None):
        """``lower_bounds[i] <= sum_j vals[i,j] * x[cols[i,j]] <= upper_bounds[i]``."""
        self.add_soft_inequality_constraints(
            cols, vals, coef_penalization=np.inf, lower_bounds=lower_bounds, upper_bounds=upper_bounds)

    def add_soft_inequality_constraints(self, cols, vals, coef_penalization, lower_bounds=None, upper_bounds=None):
        """Hard rows when the penalty is +inf, otherwise hinge penalties through auxiliary variables."""
        if np.all(coef_penalization == np.inf):
            self.add_inequality_constraints_sparse(
                crd_matrix(cols, vals), lower_bounds=lower_bounds, upper_bounds=upper_bounds)
            return None
        if np.any(coef_penalization == np.inf):
            raise NotImplementedError("mixing finite and infinite penalties is not handled")
        cols, vals = np.broadcast_arrays(cols, vals)
        aux = self.add_variables_array((cols.shape[0],), upper_bounds=None, lower_bounds=0, costs=coef_penalization)
        cols_aux = np.column_stack((cols, aux))
        if upper_bounds is None and lower_bounds is None:
            raise ValueError("a soft constraint needs at least one bound")
        if upper_bounds is not None:
            self.add_inequality_constraints(
                cols_aux, np.column_stack((vals, -np.ones((vals.shape[0], 1)))),
                lower_bounds=None, upper_bounds=upper_bounds)
        if lower_bounds is not None:
            self.add_inequality_constraints(
                cols_aux, np.column_stack((vals, np.ones((vals.shape[0], 1)))),
                lower_bounds=lower_bounds, upper_bounds=None)
        return aux

    # -- transformations ------------------------------------------------------------------------
    def remove_fixed_variables(self):
        """Drop variables with ``ub <= lb`` (reference :632-674).

        Returns ``(m_change, shift)``: ``m_change`` scatters the reduced vector back to
        full length, ``shift`` holds the fixed values.  Right-hand sides are moved by
        ``A @ shift``.
        """
        free = self.upper_bounds > self.lower_bounds
        free_ids = np.flatnonzero(free)
        nb_free = free_ids.size
        m_change = sp.coo_matrix(
            (np.ones(nb_free), (free_ids, np.arange(nb_free))), (self.nb_variables, nb_free))
        shift = np.zeros(self.nb_variables)
        shift[~free] = self.lower_bounds[~free]
        a_eq, a_ineq = self.a_equalities, self.a_inequalities
        self.b_equalities = self.b_equalities - a_eq * shift
        moved = a_ineq * shift
        if self.b_lower is not None:
            self.b_lower = self.b_lower - moved
        if self.b_upper is not None:
            self.b_upper = self.b_upper - moved
        self.costsvector = self.costsvector[free]
        if nb_free != self.nb_variables:
            a_ineq = a_ineq[:, free]
            a_eq = a_eq[:, free]
        self.nb_variables = nb_free
        self.a_inequalities = a_ineq
        self.a_equalities = a_eq
        self.lower_bounds = self.lower_bounds[free]
        self.upper_bounds = self.upper_bounds[free]
        self.is_integer = self.is_integer[free]
        return m_change, shift

    def convert_to_one_sided_inequality_system(self):
        """Rewrite ``b_lower <= A x <= b_upper`` as ``A' x <= b'`` (reference :835-879).

        Row order of ``A'``: rows with a finite upper bound in their original order,
        then the negation of rows with a finite lower bound in their original order.
        ``b_lower`` becomes ``None``.
        """
        if self.b_lower is None:
            return
        a = self.a_inequalities
        up = np.flatnonzero(self.b_upper != np.inf)
        lo = np.flatnonzero(self.b_lower != -np.inf)
        if up.size and lo.size:
            n_up = up.size
            up_pos = np.concatenate(([0], np.cumsum(self.b_upper != np.inf)))
            lo_pos = np.concatenate(([0], np.cumsum(self.b_lower != np.inf)))
            names = [{"name": d["name"], "start": up_pos[d["start"]], "end": up_pos[d["end"]]}
                     for d in self.inequalityConstraintNames]
            names += [{"name": d["name"], "start": n_up + lo_pos[d["start"]], "end": n_up + lo_pos[d["end"]]}
                      for d in self.inequalityConstraintNames]
            self.inequalityConstraintNames = names
            a = sp.vstack((a[up, :], -a[lo, :])).tocsr()
        elif lo.size:
            a = -a
        self.a_inequalities = a
        self.b_upper = np.concatenate((self.b_upper[up], -self.b_lower[lo]))
        self.b_lower = None

    def convert_to_all_equalities(self):
        """``A_ineq x - s = 0`` with one slack ``b_lower <= s <= b_upper`` per inequality row (reference :819-833);
        the solution of the original LP is the leading part of x."""
        m = self.nb_inequality_constraints()
        if m == 0:
            return
        a = self.a_inequalities
        lower = self.b_lower if self.b_lower is not None else np.full(m, -np.inf)
        self.add_variables_array(m, lower, self.b_upper)
        self._ineq.replace(None)
        self.add_inequality_constraints_sparse(sp.hstack((a, -sp.eye(m))).tocsr(), 0, 0)  # equal scalars: equalities
        self.b_lower = np.empty(0, dtype=np.float64)
        self.b_upper = np.empty(0, dtype=np.float64)
        self.inequalityConstraintNames = []

    def convert_to_all_inequalities(self):
        """Equality rows become two-sided inequality rows ``b <= A_eq x <= b``, in front of the existing ones
        (reference :881-911)."""
        m_eq = self.nb_equality_constraints()
        if m_eq == 0:
            return
        m_in = self.nb_inequality_constraints()
        lower = self.b_lower if self.b_lower is not None else np.full(m_in, -np.inf)
        upper = self.b_upper if self.b_upper is not None else np.full(m_in, np.inf)
        self.inequalityConstraintNames = list(self.equalityConstraintNames) + [
            {"name": d["name"], "start": m_eq + d["start"], "end": m_eq + d["end"]} for d in self.inequalityConstraintNames]
        self.equalityConstraintNames = []
        self.a_inequalities = sp.vstack((self.a_equalities, self.a_inequalities)).tocsr()
        self.b_lower = np.concatenate((self.b_equalities, lower))
        self.b_upper = np.concatenate((self.b_equalities, upper))
        self._eq.replace(None)
        self.b_equalities = np.empty(0, dtype=np.float64)

    def convert_to_all_inequalities_without_bounds(self):
        """Also the variable bounds become rows (one per variable with a finite bound), the variables are left free
        (reference :913-928)."""
        self.convert_to_all_inequalities()
        bounded = np.flatnonzero(~(np.isinf(self.lower_bounds) & np.isinf(self.upper_bounds)))
        m_in = self.nb_inequality_constraints()
        lower = self.b_lower if self.b_lower is not None else np.full(m_in, -np.inf)
        pick = sp.csr_matrix((np.ones(bounded.size), (np.arange(bounded.size), bounded)),
                             shape=(bounded.size, self.nb_variables))
        self.a_inequalities = sp.vstack((self.a_inequalities, pick)).tocsr()
        self.b_lower = np.concatenate((lower, self.lower_bounds[bounded]))
        self.b_upper = np.concatenate((self.b_upper, self.upper_bounds[bounded]))
        self.lower_bounds.fill(-np.inf)
        self.upper_bounds.fill(np.inf)

    def _fixed_variable_curve_terms(self, reduced, shift, ground_truth, ground_truth_indices):
        """What the eliminated variables contribute to the per-callback curves of ``solve`` (reference :1064-1093)
        when x stays on the device.  The curves are evaluated on the FULL LP at ``x_full = m_change x - shift``
        (the reference's sign, :1259): an eliminated entry of ``x_full`` is the constant ``-shift``.

        * bounds (:188-189): ``max(lb - x_full, x_full - ub)`` over the eliminated entries — a constant;
        * rows (:190-197): with ``A_full x_full = A_free x - A shift`` and the reduced right-hand sides
          ``b' = b - A shift`` (:646-656), the full residual of a row is the reduced one minus ``2 (A shift)_i`` —
          a per-row constant, in the row order of the one-sided system the solver builds
          (finite uppers, then negated finite lowers: ``+2 (A shift)`` / ``-2 (A shift)``);
        * ground truth (:1074-1082): eliminated entries contribute constant terms to the two sums.
        """
        free = self.upper_bounds > self.lower_bounds
        x_fixed = -shift[~free]
        out = {"bound_violation": max(float(np.max(self.lower_bounds[~free] - x_fixed)),
                                      float(np.max(x_fixed - self.upper_bounds[~free])))}
        offs = []
        if self.nb_equality_constraints() > 0:
            offs.append(2.0 * (self.a_equalities * shift))
        if self.nb_inequality_constraints() > 0:
            moved = 2.0 * (self.a_inequalities * shift)
            if reduced.b_lower is None:
                offs.append(moved)
            else:
                up = np.flatnonzero(np.asarray(reduced.b_upper) != np.inf)
                lo = np.flatnonzero(np.asarray(reduced.b_lower) != -np.inf)
                offs.append(np.concatenate((moved[up], -moved[lo])))
        out["row_offsets"] = np.concatenate(offs) if offs else np.zeros(0)
        if ground_truth is not None:
            gt = np.asarray(ground_truth, dtype=np.float64).ravel()
            idx = np.asarray(ground_truth_indices).ravel()
            reduced_id = np.cumsum(free) - 1
            is_free = free[idx]
            out["gt_free_reduced_ids"] = reduced_id[idx[is_free]].astype(np.int32)
            out["gt_free_values"] = gt[is_free]
            xf = -shift[idx[~is_free]]
            out["gt_fixed_sum"] = float(np.sum(np.abs(gt[~is_free] - xf)))
            out["gt_fixed_sum_rounded"] = float(np.sum(np.abs(gt[~is_free] - np.round(xf))))
        return out

    # -- solve ------------------------------------------------------------------------------------
    def solve(
        self,
        method="admm",
        get_timing=True,
        x0=None,
        nb_iter=10000,
        max_time=None,
        callback_func=None,
        nb_iter_plot=10,
        plot_solution=None,
        ground_truth=None,
        ground_truth_indices=None,
        **solver_options,
    ):
        """Reference ``SparseLP.solve`` (:990-1002) restricted to the CP-PPD branch (:1243-1288).

        The per-callback curve attributes (``itrn_curve``, ``pobj_curve``, ``dobj_curve``,
        ``distance_to_ground_truth`` ...) are filled exactly as the reference does
        (:1018-1028, :1064-1093).  As in the reference, the ``callback_func`` argument
        is shadowed by the internal curve recorder and never called; ``plot_solution``
        is the user hook.  ``solver_options`` are forwarded to ``chambolle_pock_ppd`` as
        extra keyword-only arguments (``device=``, ``flags=`` ...).  When no variable is fixed and
        ``plot_solution`` is None the curves are evaluated on the device (``device_curves=False`` forces
        the reference's host-side evaluation, which downloads x at every callback).
        """
        if method not in solving_methods:
            if method in _reference_only_methods:
                raise NotImplementedError(
                    "method %r belongs to the reference package but is outside this build's scope; "
                    "only 'chambolle_pock_ppd' is provided" % method)
            raise ValueError("unknown LP solver method %r; available: %s" % (method, solving_methods))
        start = time.perf_counter()
        self.distance_to_ground_truth = []
        self.distanceToGroundTruthAfterRounding = []
        self.opttime_curve = []
        self.dopttime_curve = []
        self.pobj_curve = []
        self.dobj_curve = []
        self.pobjbound = []
        self.max_violated_inequality = []
        self.max_violated_equality = []
        self.max_violated_constraint = []
        self.itrn_curve = []

        def record(niter, solution, energy1, energy2, duration, max_violated_equality,
                   max_violated_inequality, is_active_variable=None):
            if ground_truth is not None:
                picked = solution[ground_truth_indices]
                self.distance_to_ground_truth.append(np.mean(np.abs(ground_truth - picked)))
                self.distanceToGroundTruthAfterRounding.append(
                    np.mean(np.abs(ground_truth - np.round(picked))))
            self.itrn_curve.append(niter)
            self.opttime_curve.append(duration)
            self.dopttime_curve.append(duration)
            self.dobj_curve.append(energy2)
            self.pobj_curve.append(energy1)
            self.max_violated_constraint.append(self.max_constraint_violation(solution))
            self.max_violated_equality.append(max_violated_equality)
            self.max_violated_inequality.append(max_violated_inequality)
            if plot_solution is not None:
                plot_solution(niter, solution, is_active_variable=is_active_variable)

        reduced = copy.deepcopy(self)
        m_change, shift = reduced.remove_fixed_variables()
        m_change = m_change.tocsr()

        def to_full(v):
            # NOTE: the reference maps back with "- shift" (:1259, :1288) although the change of
            # variables is x_full = M y + shift; reproduced as-is for parity.
            return m_change * v - shift

        def reduced_callback(niter, solution, energy1, energy2, duration, mv_eq, mv_ineq):
            record(niter, to_full(solution), energy1, energy2, duration, mv_eq, mv_ineq)

        solver_args = (reduced.costsvector, reduced.a_equalities, reduced.b_equalities, reduced.a_inequalities,
                       reduced.b_lower, reduced.b_upper, reduced.lower_bounds, reduced.upper_bounds)
        want_device_curves = solver_options.pop("device_curves", True)
        device_curves = want_device_curves and plot_solution is None
        if int(solver_options.get("n_gpus") or 1) > 1:
            # several GPUs from this one process (pysparselp_b200/multi_gpu.py): the solver state lives in helper
            # processes too, so the curves come through the callback, evaluated on the host like the reference does
            device_curves = False
        fixed = None
        if device_curves and reduced.nb_variables != self.nb_variables:
            fixed = self._fixed_variable_curve_terms(reduced, shift, ground_truth, ground_truth_indices)
        if device_curves:
            # Nobody asked to see x: every per-callback curve of the reference (:1074-1091) is evaluated on the
            # device inside the stats block and x never leaves the GPU before the end.  Values equal the host path
            # up to the summation order of the two distance means.  When variables were eliminated (:632-674) the
            # curves are those of the FULL LP at the mapped-back point m_change * x - shift: the eliminated entries
            # contribute constants (bounds, ground truth) and shift every row residual by a constant (`fixed`).
            solve_kw = {k: solver_options[k] for k in ("device", "flags", "distributed", "partition_granule", "kernel_variant",
                                                          "long_row_threshold")
                        if k in solver_options}
            solver = make_solver(*solver_args, x0=None, alpha=1, theta=1, **solve_kw)
            best_integer_solution = None
            if solver is None:
                device_curves = False
            else:
                try:
                    gt_total = 0 if ground_truth is None else np.size(ground_truth)
                    if ground_truth is not None and fixed is None:
                        solver.set_ground_truth(ground_truth_indices, ground_truth)
                    elif ground_truth is not None and fixed["gt_free_values"].size:
                        solver.set_ground_truth(fixed["gt_free_reduced_ids"], fixed["gt_free_values"])
                    if fixed is not None:
                        solver.set_row_offsets(fixed["row_offsets"])

                    def record_stats(niter, st, duration):
                        if ground_truth is not None and fixed is None:
                            self.distance_to_ground_truth.append(st["distance_to_ground_truth"])
                            self.distanceToGroundTruthAfterRounding.append(st["distance_to_ground_truth_rounded"])
                        elif ground_truth is not None:
                            # the device means run over the free entries only: back to sums, add the constants
                            nfree = fixed["gt_free_values"].size
                            self.distance_to_ground_truth.append(
                                (st["distance_to_ground_truth"] * nfree + fixed["gt_fixed_sum"]) / gt_total)
                            self.distanceToGroundTruthAfterRounding.append(
                                (st["distance_to_ground_truth_rounded"] * nfree + fixed["gt_fixed_sum_rounded"]) / gt_total)
                        self.itrn_curve.append(niter)
                        self.opttime_curve.append(duration)
                        self.dopttime_curve.append(duration)
                        self.dobj_curve.append(st["energy2"])
                        self.pobj_curve.append(st["energy1"])
                        # max_constraint_violation (:186-204) from the device maxima, folded in its order
                        worst = max(0, st["max_bound_violation"])
                        if fixed is not None:
                            worst = max(worst, fixed["bound_violation"])
                        if self.nb_equality_constraints() > 0:
                            worst = max(worst, st["max_violated_equality_full"])
                        if self.nb_inequality_constraints() > 0:
                            worst = max(worst, st["max_violated_inequality_full"])
                        self.max_violated_constraint.append(worst)
                        self.max_violated_equality.append(st["max_violated_equality"])
                        self.max_violated_inequality.append(st["max_violated_inequality"])

                    x, best_integer_solution = run_schedule(solver, nb_iter, None, max_time, False, nb_iter_plot,
                                                            bool(solver_options.get("verbose", False)), start,
                                                            stats_func=record_stats)
                finally:
                    solver.close()
        if not device_curves:
            x, best_integer_solution = chambolle_pock_ppd(
                *solver_args,
                x0=None,
                alpha=1,
                theta=1,
                nb_max_iter=nb_iter,
                callback_func=reduced_callback,
                max_time=max_time,
                save_problem=False,
                nb_iter_plot=nb_iter_plot,
                **solver_options,
            )
        x = to_full(x)
        self.best_integer_solution = None if best_integer_solution is None else to_full(best_integer_solution)
        elapsed = time.perf_counter() - start
        return (x, elapsed) if get_timing else x
