"""ctypes binding of ``libcpppd.so`` (C ABI declared in ``include/cpppd.h``).

The library is built in-tree by ``pysparselp_b200/build.py`` (``nvcc`` for sm_100a) and
is the ONLY compute path: if it is missing or no CUDA device is present, loading /
``cpppd_create`` fails loudly — there is no CPU fallback in the product.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcpppd.so")

ABI_VERSION = 7
KERNEL_VARIANTS = 9

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)

FLAG_VALUE_DICT = 1 << 0
FLAG_CONST_VECTORS = 1 << 1
FLAG_NO_GRAPH = 1 << 2
FLAG_REORDER = 1 << 3
FLAG_GRAPH_COMM = 1 << 4
FLAG_NO_P2P = 1 << 5
FLAG_NO_REORDER = 1 << 6
FLAG_FUSED_HALO = 1 << 7
FLAG_NO_AUTOTUNE = 1 << 8
FLAG_TINY_PERSISTENT = 1 << 9
FLAG_BANDED = 1 << 10
FLAG_NO_BANDED = 1 << 11
FLAG_NO_TINY_PERSISTENT = 1 << 12
FLAG_NO_DENSE_HALO = 1 << 13

VEC_X, VEC_XBAR, VEC_Y, VEC_T, VEC_SIGMA, VEC_BEST_INTEGER, VEC_D = range(7)


class Problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("n", C.c_int64), ("m_eq", C.c_int64), ("m_ineq", C.c_int64), ("nnz", C.c_int64),
        ("indptr", C.c_void_p), ("indices", C.c_void_p), ("values", C.c_void_p),
        ("indptr_bits", C.c_int32), ("index_bits", C.c_int32),
        ("c", C.c_void_p), ("b", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p), ("x0", C.c_void_p),
        ("alpha", C.c_double), ("theta", C.c_double), ("one_plus_theta", C.c_double),
        ("stream", C.c_void_p), ("flags", C.c_uint32), ("kernel_variant", C.c_int32),
        ("alloc", ALLOC_FN), ("free", FREE_FN), ("alloc_user", C.c_void_p),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("comm_id", C.c_void_p),
        ("partition_granule", C.c_int64), ("comm", C.c_void_p), ("long_row_threshold", C.c_int64),
        ("band_window", C.c_int64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("niter", C.c_int64), ("energy1", C.c_double), ("energy2", C.c_double),
        ("max_violated_equality", C.c_double), ("max_violated_inequality", C.c_double),
        ("energy_rounded", C.c_double), ("max_violated_equality_rounded", C.c_double),
        ("best_integer_energy", C.c_double), ("frac_zero_xbar", C.c_double),
        ("max_bound_violation", C.c_double), ("distance_to_ground_truth", C.c_double),
        ("distance_to_ground_truth_rounded", C.c_double),
        ("max_violated_equality_full", C.c_double), ("max_violated_inequality_full", C.c_double),
        ("feasible", C.c_int32), ("improved", C.c_int32), ("have_best_integer", C.c_int32),
        ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if name != "reserved"}


class Info(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("m_eq", C.c_int64), ("m_ineq", C.c_int64), ("nnz", C.c_int64),
        ("a_padded_entries", C.c_int64), ("at_padded_entries", C.c_int64), ("device_bytes", C.c_int64),
        ("bytes_per_iteration_algorithmic", C.c_int64), ("bytes_per_iteration_actual", C.c_int64),
        ("value_bytes", C.c_int32), ("const_vector_mask", C.c_int32), ("sm_count", C.c_int32),
        ("world_size", C.c_int32), ("rank", C.c_int32), ("primal_variant", C.c_int32),
        ("n_local", C.c_int64), ("m_local", C.c_int64), ("m_eq_local", C.c_int64),
        ("n_ghost", C.c_int64), ("m_ghost", C.c_int64),
        ("nnz_local_rows", C.c_int64), ("nnz_local_cols", C.c_int64),
        ("halo_send_bytes_per_iteration", C.c_int64), ("partition_granule", C.c_int64),
        ("dual_variant", C.c_int32), ("autotuned", C.c_int32), ("variant_ms", (C.c_float * KERNEL_VARIANTS) * 2),
        ("long_rows", C.c_int64), ("long_cols", C.c_int64), ("long_entries", C.c_int64),
        ("balanced_split", C.c_int32), ("tiny_persistent", C.c_int32),
        ("band_windows", C.c_int32 * 2), ("band_in_use", C.c_int32 * 2), ("band_ms", C.c_float * 2),
        ("band_sectors_per_gather", C.c_float * 2), ("band_window_bytes", C.c_int64),
        ("dense_halo", C.c_int32), ("reserved2", C.c_int32),
        ("band_shape", C.c_int32 * 2), ("band_shape_ms", (C.c_float * 8) * 2),
    ]

    def as_dict(self):
        arrays = ("band_windows", "band_in_use", "band_ms", "band_sectors_per_gather", "band_shape")
        out = {name: getattr(self, name) for name, _ in self._fields_
               if name not in ("variant_ms", "band_shape_ms", "reserved", "reserved2") + arrays}
        out["band_shape_ms"] = [list(self.band_shape_ms[0]), list(self.band_shape_ms[1])]
        for name in arrays:  # [A (dual half), A^T (primal half)]
            out[name] = list(getattr(self, name))
        out["variant_ms"] = {"k_primal": list(self.variant_ms[0]), "k_dual": list(self.variant_ms[1])}
        return out


# every symbol include/cpppd.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cpppd_abi_version": (C.c_int, []),
    "cpppd_comm_unique_id": (C.c_int, [C.c_void_p]),
    "cpppd_comm_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "cpppd_comm_destroy": (C.c_int, [C.c_void_p]),
    "cpppd_create": (C.c_int, [C.POINTER(Problem), C.POINTER(C.c_void_p)]),
    "cpppd_destroy": (C.c_int, [C.c_void_p]),
    "cpppd_last_error": (C.c_char_p, [C.c_void_p]),
    "cpppd_iterate": (C.c_int, [C.c_void_p, C.c_int64]),
    "cpppd_primal_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "cpppd_stats_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "cpppd_dual_step": (C.c_int, [C.c_void_p]),
    "cpppd_sync": (C.c_int, [C.c_void_p]),
    "cpppd_read_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "cpppd_time_iterations": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float)]),
    "cpppd_time_kernels": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cpppd_get_vector": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cpppd_set_vector": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "cpppd_get_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "cpppd_set_ground_truth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "cpppd_set_row_offsets": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cpppd_get_layout": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
    "cpppd_iteration_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


class CpppdError(RuntimeError):
    """A libcpppd entry point returned a negative status."""

    def __init__(self, code, message):
        super().__init__("libcpppd error %d: %s" % (code, message))
        self.code = code


def load_library(path=None):
    """dlopen the in-tree library and type every symbol; raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise RuntimeError(
            "%s is not built — run `python -m pysparselp_b200.build` (needs nvcc). "
            "This package has no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cpppd_abi_version() != ABI_VERSION:
        raise RuntimeError("libcpppd ABI %d != binding ABI %d: rebuild" % (lib.cpppd_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(lib, handle, code):
    if code < 0:
        msg = lib.cpppd_last_error(handle)
        raise CpppdError(code, msg.decode("utf-8", "replace") if msg else "")
    return code
