"""CPU, build container only: the MPS parser of this package against the reference's (run live) on random
fixed-column MPS files: N / E / L / G rows, one or two entries per COLUMNS / RHS line, every bound type the reference
accepts (UP, LO, FX, FR, MI, PL), comment lines.  Skipped where /root/reference does not exist."""
import importlib
import io

import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def field_line(kind="", name1="", name2="", num1="", name3="", num2=""):
    """Fixed MPS columns: 2-3 type, 5-12 name, 15-22 name, 25-36 number, 40-47 name, 50-61 number."""
    # (the reference slices the fields by column and looks names up unstripped: a line must reach column 22, and one
    # with a single entry must end before column 40)
    if not name3:
        return " %-2s %-8s  %-8s  %12s\n" % (kind, name1, name2, num1)
    return " %-2s %-8s  %-8s  %12s   %-8s  %12s\n" % (kind, name1, name2, num1, name3, num2)


def random_mps(seed):
    rng = np.random.default_rng(seed)
    n_rows, n_cols = int(rng.integers(1, 12)), int(rng.integers(1, 10))
    rows = [("N", "COST")] + [(str(rng.choice(["E", "L", "G"])), "R%d" % i) for i in range(n_rows)]
    out = ["NAME          FUZZ%d\n" % seed, "ROWS\n"]
    out += [field_line(kind, name) for kind, name in rows]
    out.append("COLUMNS\n")
    num = lambda: "%.3f" % (np.round(rng.standard_normal() * 5, 3))  # noqa: E731
    cols = ["X%d" % j for j in range(n_cols)]
    for col in cols:
        touched = [rows[i][1] for i in sorted(rng.choice(len(rows), size=int(rng.integers(1, len(rows) + 1)), replace=False))]
        if rng.random() < 0.2:
            out.append("* a comment line\n")
        while touched:
            if len(touched) >= 2 and rng.random() < 0.6:
                out.append(field_line("", col, touched[0], num(), touched[1], num()))
                touched = touched[2:]
            else:
                out.append(field_line("", col, touched[0], num()))
                touched = touched[1:]
    out.append("RHS\n")
    with_rhs = [name for kind, name in rows[1:] if rng.random() < 0.7]
    while with_rhs:
        if len(with_rhs) >= 2 and rng.random() < 0.5:
            out.append(field_line("", "RHS", with_rhs[0], num(), with_rhs[1], num()))
            with_rhs = with_rhs[2:]
        else:
            out.append(field_line("", "RHS", with_rhs[0], num()))
            with_rhs = with_rhs[1:]
    if rng.random() < 0.8:
        out.append("BOUNDS\n")
        for col in cols:
            if rng.random() < 0.6:
                kind = str(rng.choice(["UP", "LO", "FX", "FR", "MI", "PL"]))
                out.append(field_line(kind, "BND", col, num() if kind in ("UP", "LO", "FX") else ""))
    out.append("ENDATA\n")
    return "".join(out)


@pytest.mark.parametrize("seed", range(40))
def test_mps_parser_equals_the_live_reference(seed):
    from pysparselp_b200.MPSparser import mps_parser

    ref_loader.load_reference()
    ref_parser = importlib.import_module("pysparselp.MPSparser").mps_parser
    text = random_mps(seed)
    want, got = ref_parser(io.StringIO(text)), mps_parser(io.StringIO(text))
    for key, w in want.items():
        g = got[key]
        if hasattr(w, "tocsr"):
            w, g = w.tocsr(), g.tocsr()
            assert w.shape == g.shape and np.array_equal(w.indptr, g.indptr), (key, text)
            assert np.array_equal(w.indices, g.indices) and np.array_equal(w.data, g.data), (key, text)
        elif isinstance(w, np.ndarray):
            assert np.array_equal(w, g), (key, text)
        elif isinstance(w, str):
            assert w.strip() == g.strip() or key == "problem_name", (key, w, g)  # the reference slices a blank name
        else:
            assert w == g, (key, text)
