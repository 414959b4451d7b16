"""pysparselp_b200 — B200-native Chambolle-Pock PPD LP solver behind PySparseLP's API.

Public surface (mirrors the reference package for the CP-PPD path):

* ``pysparselp_b200.ChambollePockPPD.chambolle_pock_ppd``  — the solver entry point
* ``pysparselp_b200.SparseLP.SparseLP`` / ``solving_methods`` — modeling layer + ``solve()``
* ``pysparselp_b200.MPSparser.mps_parser``, ``pysparselp_b200.netlib.get_problem``
* ``pysparselp_b200.examples`` — Potts segmentation and L1-SVM problem builders
* ``pysparselp_b200.generators`` — direct CSR emitters for the large benchmark LPs

Importing the package does not touch CUDA; the compiled core (``csrc/libcpppd.so``) is
loaded on first solve and there is no CPU fallback.
"""
__version__ = "0.1.0"
