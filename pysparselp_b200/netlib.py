"""Locate a vendored netlib problem and parse it (reference ``pysparselp/netlib.py:35-73``).

The reference falls back to an FTP/HTTP download when the files are missing; this
build has no network path — only the problems vendored under ``data/`` are served.
"""
import os

from .MPSparser import mps_parser

_HERE = os.path.dirname(os.path.abspath(__file__))


def get_problem(problem_name, data_dir=None):
    """``data_dir``: a folder laid out like the reference's ``pysparselp/data`` (``netlib/NAME.SIF``,
    ``perPlex/name.txt``) for problems that are not vendored here; default: this package's ``data/``."""
    root = os.path.join(_HERE, "data") if data_dir is None else data_dir
    lp_file = os.path.join(root, "netlib", problem_name.upper() + ".SIF")
    sol_file = os.path.join(root, "perPlex", problem_name.lower() + ".txt")
    if not os.path.isfile(lp_file):
        raise FileNotFoundError(
            "netlib problem %s is not vendored (%s); downloading is not supported" % (problem_name, lp_file))
    with open(lp_file, "r") as f:
        if os.path.isfile(sol_file):
            with open(sol_file, "r") as fs:
                return mps_parser(f, fs)
        return mps_parser(f, None)
