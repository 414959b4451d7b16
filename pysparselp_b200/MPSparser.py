"""Fixed-column MPS reader + perPlex exact-solution reader (host side, one-time text parse).

Mirrors the interface of the reference's ``pysparselp/MPSparser.py:10-271``
(``mps_parser(f, fsol=None)`` returning a dict with the same keys) so that
``tests/test_netlib.py``-style drivers run unchanged.  Written as a small
section-driven state machine; supported sections are NAME, ROWS, COLUMNS, RHS,
BOUNDS (RANGES and integer bound types raise, as in the reference ``:71-73,:175-177``).

Conventions kept from the reference because they change the LP that reaches the solver:

* ``G`` rows get ``b_lower = 0, b_upper = +inf``; ``L`` rows ``b_lower = -inf, b_upper = 0``;
  ``E`` rows default to ``b_eq = 0`` (``:83-99``) until the RHS section overrides them.
* variables default to ``[0, +inf)`` and cost 0 (``:108-116``); ``UP`` never touches the
  lower bound, ``MI`` only lowers it, ``FR`` frees both sides (``:160-174``).
* inequality rows are numbered in order of appearance among L/G rows, equality rows
  among E rows, variables in order of first appearance in COLUMNS.
"""
import numpy as np
from scipy import sparse

# field windows of the fixed MPS layout (0-based, end-exclusive):
#   type 2-3 | name 5-12 | name 15-22 | number 25-36 | name 40-47 | number 50-61
_WINDOWS = ((1, 3), (4, 12), (14, 22), (24, 36), (39, 47), (49, 61))
_SECTIONS = ("NAME", "ROWS", "COLUMNS", "RHS", "RANGES", "BOUNDS", "ENDATA")


def _fields(line):
    return [line[a:b].strip() for a, b in _WINDOWS]


def _pairs(fields):
    """(row name, value) pairs carried by a COLUMNS/RHS data line."""
    for k in (2, 4):
        if fields[k] == "":
            return
        yield fields[k], float(fields[k + 1])


def mps_parser(f, fsol=None):
    """Parse an MPS stream ``f`` (and optionally a perPlex solution stream ``fsol``)."""
    row_kind, row_pos = {}, {}
    var_pos, var_cost, var_lo, var_up = {}, [], [], []
    ineq_lo, ineq_up, eq_rhs = [], [], []
    entries_ineq, entries_eq = [], []
    problem_name, costname = None, None
    section = None

    for raw in f:
        line = raw.rstrip("\n")
        if not line.strip() or line.startswith("*"):
            continue
        head = line.split()[0]
        if not line[0].isspace() and head in _SECTIONS:
            if head == "ENDATA":
                break
            if head == "RANGES":
                raise NotImplementedError("RANGES section is not supported")
            if head == "NAME":
                problem_name = line[14:22].strip() or (line.split()[1] if len(line.split()) > 1 else "")
                continue
            section = head
            continue
        t = _fields(line)
        if section == "ROWS":
            kind, name = t[0], t[1]
            if name in row_kind:
                raise ValueError("row %r declared twice" % name)
            row_kind[name] = kind
            if kind == "N":
                costname = name
            elif kind == "G":
                row_pos[name] = len(ineq_lo)
                ineq_lo.append(0.0)
                ineq_up.append(np.inf)
            elif kind == "L":
                row_pos[name] = len(ineq_lo)
                ineq_lo.append(-np.inf)
                ineq_up.append(0.0)
            elif kind == "E":
                row_pos[name] = len(eq_rhs)
                eq_rhs.append(0.0)
        elif section == "COLUMNS":
            name = t[1]
            if name not in var_pos:
                var_pos[name] = len(var_cost)
                var_cost.append(0.0)
                var_lo.append(0.0)
                var_up.append(np.inf)
            j = var_pos[name]
            for rname, value in _pairs(t):
                kind = row_kind[rname]
                if kind == "N":
                    var_cost[j] = value
                elif kind == "E":
                    entries_eq.append((row_pos[rname], j, value))
                else:
                    entries_ineq.append((row_pos[rname], j, value))
        elif section == "RHS":
            for rname, value in _pairs(t):
                kind = row_kind[rname]
                if kind == "N":
                    raise ValueError("right-hand side on the objective row is not supported")
                if kind == "L":
                    ineq_up[row_pos[rname]] = value
                elif kind == "G":
                    ineq_lo[row_pos[rname]] = value
                else:
                    eq_rhs[row_pos[rname]] = value
        elif section == "BOUNDS":
            kind, vname = t[0], t[2]
            j = var_pos[vname]
            if kind == "UP":
                var_up[j] = float(t[3])
            elif kind == "LO":
                var_lo[j] = float(t[3])
            elif kind == "FR":
                var_lo[j], var_up[j] = -np.inf, np.inf
            elif kind == "FX":
                var_lo[j] = var_up[j] = float(t[3])
            elif kind == "MI":
                var_lo[j] = -np.inf
            elif kind == "PL":
                var_up[j] = np.inf
            elif kind in ("BV", "LI", "UI"):
                raise NotImplementedError("integer bound types are not supported")

    nb_var = len(var_cost)

    def assemble(entries, nrows):
        m = sparse.dok_matrix((nrows, nb_var))
        for i, j, v in entries:
            m[i, j] = v
        return m

    out = {
        "cost_vector": np.array(var_cost, dtype=np.float64),
        "upper_bounds": np.array(var_up, dtype=np.float64),
        "lower_bounds": np.array(var_lo, dtype=np.float64),
        "a_eq": assemble(entries_eq, len(eq_rhs)),
        "b_eq": np.array(eq_rhs, dtype=np.float64),
        "a_ineq": assemble(entries_ineq, len(ineq_lo)),
        "b_lower": np.array(ineq_lo, dtype=np.float64),
        "b_upper": np.array(ineq_up, dtype=np.float64),
        "problem_name": problem_name,
        "costname": costname,
        "solution": None,
    }
    if fsol is not None:
        out["solution"] = _perplex_solution(fsol, var_pos, var_lo, var_up)
    return out


def _perplex_solution(fsol, var_pos, var_lo, var_up):
    """Exact optimal vertex written by perPlex 1.00 (reference ``MPSparser.py:207-269``).

    A variable's value is its ``V Value`` rational ``p/q`` evaluated in float64
    (falling back to the decimal rendering when that is NaN), unless a later
    ``V State : on lower/upper/both`` line pins it to the parsed bound.
    """
    sol = np.full(len(var_pos), np.nan)
    part, j = None, None
    for raw in fsol:
        line = raw.rstrip("\n")
        if line.startswith("- EOF"):
            break
        if line.startswith("- Variables"):
            part = "V"
            continue
        if line.startswith("- Constraints"):
            part = "C"
            continue
        if part != "V" or not line.startswith("V "):
            continue
        key, _, rest = line.partition(":")
        key = key[2:].strip()
        rest = rest.strip()
        if key == "Name":
            j = var_pos[rest]
        elif key == "Value":
            decimal, _, rational = rest.partition("=")
            num, _, den = rational.partition("/")
            val = float(num) / float(den) if den.strip() else float(num)
            sol[j] = float(decimal) if np.isnan(val) else val
        elif key == "State":
            if rest.startswith("on lower"):
                sol[j] = var_lo[j]
            elif rest.startswith("on upper"):
                sol[j] = var_up[j]
            elif rest.startswith("on both"):
                assert var_up[j] == var_lo[j]
                sol[j] = var_up[j]
    return sol
