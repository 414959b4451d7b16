"""Several GPUs behind ONE Python process: ``chambolle_pock_ppd(..., n_gpus=N)`` / ``lp.solve(..., n_gpus=N)``.

The reference call site (``pysparselp/SparseLP.py:1270-1288``) is one function call in one process; the CUDA core scales
as one process per GPU (``torch.distributed`` rendezvous, NCCL / peer-memory halos).  This module
bridges the two: the calling process becomes rank 0, ``N - 1`` helper processes are spawned for the other GPUs, the
LP travels to them once through POSIX shared memory (no pickling of gigabytes), every rank runs the ordinary
distributed solve, and the helpers exit when the call returns.  Callbacks run in the calling process only (rank 0
fetches x; the helpers take part in the collective fetch), ``max_time`` is decided by rank 0 for all.

Cost: helper start-up (interpreter + torch import + CUDA context + NCCL communicator) is 10-20 s per call — meant for
LPs whose solve takes minutes on one GPU, not for the small test LPs.
"""
import os
import pickle
import socket
import subprocess
import sys
import tempfile
from multiprocessing import shared_memory

import numpy as np
import scipy.sparse as sp

_ARRAY_ARGS = ("c", "beq", "b_lower", "b_upper", "lb", "ub", "x0")
_MATRIX_ARGS = ("a_eq", "a_ineq")


def _free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def pack_lp(lp_args):
    """dict of the solver's array / matrix arguments -> (spec, shared memory blocks).  ``spec`` is small and picklable:
    per argument None, ("array", block name, dtype, shape) or ("csr", shape, data / indices / indptr specs)."""
    blocks = []

    def put(a):
        a = np.ascontiguousarray(a)
        shm = shared_memory.SharedMemory(create=True, size=max(a.nbytes, 1))
        np.ndarray(a.shape, dtype=a.dtype, buffer=shm.buf)[...] = a
        blocks.append(shm)
        return ("array", shm.name, a.dtype.str, a.shape)

    spec = {}
    for name in _ARRAY_ARGS:
        v = lp_args.get(name)
        spec[name] = None if v is None else put(np.asarray(v, dtype=np.float64).ravel())
    for name in _MATRIX_ARGS:
        m = lp_args.get(name)
        if m is None:
            spec[name] = None
            continue
        m = m if sp.isspmatrix_csr(m) else sp.csr_matrix(m)
        spec[name] = ("csr", m.shape, put(m.data), put(m.indices), put(m.indptr))
    return spec, blocks


def unpack_lp(spec, foreign=False):
    """Inverse of pack_lp: arrays are views of the shared blocks (kept alive in the second result).  `foreign`: this
    process did not create the blocks (a helper)."""
    keep = []

    def get(item):
        _, name, dtype, shape = item
        shm = shared_memory.SharedMemory(name=name)
        if foreign:  # attaching registers the block with this process's resource tracker, which would unlink it (and
            # warn about a leak) when this process exits; the block belongs to the process that created it
            try:
                from multiprocessing import resource_tracker

                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        keep.append(shm)
        return np.ndarray(shape, dtype=np.dtype(dtype), buffer=shm.buf)

    out = {}
    for name in _ARRAY_ARGS:
        out[name] = None if spec[name] is None else get(spec[name])
    for name in _MATRIX_ARGS:
        item = spec[name]
        if item is None:
            out[name] = None
        else:
            _, shape, data, indices, indptr = item
            m = sp.csr_matrix(shape, dtype=np.float64)
            m.data, m.indices, m.indptr = get(data), get(indices), get(indptr)  # (no validation pass, no copy)
            out[name] = m
    return out, keep


def _helper(rank, world, port, spec, solve_kw, devices):
    """Body of a helper process: rank `rank` of the distributed solve; its result is discarded."""
    import torch
    import torch.distributed as dist

    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    torch.cuda.set_device(devices[rank])
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", devices[rank]))
    lp, keep = unpack_lp(spec, foreign=True)
    solve_kw = dict(solve_kw, save_problem=False, verbose=False)
    try:
        chambolle_pock_ppd(lp["c"], lp["a_eq"], lp["beq"], lp["a_ineq"], lp["b_lower"], lp["b_upper"], lp["lb"], lp["ub"],
                           x0=lp["x0"], callback_func=None, distributed=True, device=devices[rank], **solve_kw)
    finally:
        _leave_group(dist)
        del lp
        for shm in keep:
            shm.close()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)  # (no interpreter teardown: nothing of this process is needed any more, and a communicator left to an
    #               atexit hook would wait for peers that are not exiting)


def _leave_group(dist):
    """Every rank of the call destroys its NCCL communicator of libcpppd at the same point (ncclCommDestroy waits for
    the peers), then the torch process group."""
    from pysparselp_b200.ChambollePockPPD import _destroy_cached_comms

    dist.barrier()
    _destroy_cached_comms()
    dist.barrier()
    dist.destroy_process_group()


def solve_on_gpus(n_gpus, lp_args, solve_kw, callback_func=None, devices=None):
    """Run ``chambolle_pock_ppd`` on `n_gpus` GPUs of this node from the calling process; returns what it returns."""
    import torch
    import torch.distributed as dist

    from .ChambollePockPPD import chambolle_pock_ppd

    n_gpus = int(n_gpus)
    if n_gpus < 2:
        raise ValueError("solve_on_gpus needs n_gpus >= 2")
    if not torch.cuda.is_available() or torch.cuda.device_count() < n_gpus:
        raise RuntimeError("n_gpus=%d but %d CUDA devices are visible" % (
            n_gpus, torch.cuda.device_count() if torch.cuda.is_available() else 0))
    if dist.is_available() and dist.is_initialized():
        raise RuntimeError("n_gpus= spawns its own process group; this process already has one "
                           "(under torchrun, call chambolle_pock_ppd on every rank instead: distributed=None)")
    devices = list(range(n_gpus)) if devices is None else [int(d) for d in devices]
    port = _free_port()
    spec, blocks = pack_lp(lp_args)
    # helpers are fresh interpreters running this module as a script (not multiprocessing's spawn / fork: no re-import
    # of the caller's __main__, no CUDA context inherited through fork); their orders travel in a small pickle
    order = tempfile.NamedTemporaryFile(prefix="cpppd_multi_", suffix=".pkl", delete=False)
    pickle.dump({"world": n_gpus, "port": port, "spec": spec, "solve_kw": solve_kw, "devices": devices}, order)
    order.close()
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] +
                                        [p for p in env.get("PYTHONPATH", "").split(os.pathsep) if p])
    helpers = []
    try:
        for r in range(1, n_gpus):
            helpers.append(subprocess.Popen([sys.executable, "-m", "pysparselp_b200.multi_gpu", order.name, str(r)], env=env))
        torch.cuda.set_device(devices[0])
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=n_gpus,
                                device_id=torch.device("cuda", devices[0]))
        try:
            a = lp_args
            out = chambolle_pock_ppd(a["c"], a["a_eq"], a["beq"], a["a_ineq"], a["b_lower"], a["b_upper"], a["lb"], a["ub"],
                                     x0=a.get("x0"), callback_func=callback_func, distributed=True, device=devices[0],
                                     **solve_kw)
        finally:
            _leave_group(dist)
        for p in helpers:
            code = p.wait(timeout=120)
            if code != 0:
                raise RuntimeError("a helper process of the multi-GPU solve exited with code %r" % (code,))
        return out
    finally:
        for p in helpers:
            if p.poll() is None:
                p.terminate()
        for shm in blocks:
            shm.close()
            try:
                shm.unlink()
            except FileNotFoundError:
                pass
        try:
            os.unlink(order.name)
        except OSError:
            pass


if __name__ == "__main__":  # helper process: python -m pysparselp_b200.multi_gpu <orders.pkl> <rank>
    with open(sys.argv[1], "rb") as _f:
        _o = pickle.load(_f)
    _helper(int(sys.argv[2]), _o["world"], _o["port"], _o["spec"], _o["solve_kw"], _o["devices"])
