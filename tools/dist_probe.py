"""Scratch: which configuration of a distributed solve reproduces the C port's digest on a random LP
(torchrun --nproc-per-node N tools/dist_probe.py [size])."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bench
from pysparselp_b200 import _cabi as F, generators
from pysparselp_b200.ChambollePockPPD import make_solver

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank = dist.get_rank()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
want = json.load(open(bench.DIGESTS))[bench.workload_name("random", size)]["sha256"]
lp, _ = bench.build_workload("random", size, pinned=False)
args = generators.lp_args(lp)
os.environ["CPPPD_BAND_WINDOW_MB"] = sys.argv[2] if len(sys.argv) > 2 else "0.25"
from oracle.c_port import COracle
from oracle import c_port
c_port.set_threads()
co = COracle(*args)
T_ref, s_ref = co.T.copy(), co.sigma.copy()
co.iterate(1)


def diff(name, got, want, owner_ids=None):
    bad = np.flatnonzero(got != want)
    return "%s bad %d/%d first %r" % (name, bad.size, want.size, bad[:4].tolist())


for label, flags, env in (("sell+nccl", F.FLAG_NO_BANDED | F.FLAG_NO_P2P, {}), ("sell", F.FLAG_NO_BANDED, {}),
                          ("banded", F.FLAG_BANDED, {})):
    os.environ.update(env)
    s = make_solver(*args, flags=flags)
    Tg, sg = s.get_preconditioners()
    own_c, ghost_c = s.layout(True)
    own_r, ghost_r = s.layout(False)
    s.iterate(1)
    x1, xb1, y1 = s.get_x(), s.get_xbar(), s.get_y()
    print("rank %d %-10s %s | %s | %s | %s | %s | own cols %d..%d (%d) rows %d..%d (%d) sorted %s %s" % (
        rank, label, diff("T", Tg, T_ref), diff("sigma", sg, s_ref), diff("x1", x1, co.x), diff("xbar1", xb1, co.xbar),
        diff("y1", y1, co.y), own_c.min(), own_c.max(), own_c.size, own_r.min(), own_r.max(), own_r.size,
        bool(np.all(np.diff(own_c) > 0)), bool(np.all(np.diff(own_r) > 0))), flush=True)
    s.close()
    s = make_solver(*args, flags=flags)
    out = []
    for its in (1, 5):
        s.iterate(its)
        x, y = s.get_x(), s.get_y()
        out.append(bench.iterate_digest(x, y)[:12])
    info = s.info()
    s.close()
    for k in env:
        os.environ.pop(k)
    ok = bench.iterate_digest(x, y) == want
    if rank == 0:
        print("%-20s %s after1 %s after6 %s band %r shape %r split %d ghosts %d/%d neg0 x %d y %d" % (
            label, "OK " if ok else "BAD", out[0], out[1], info["band_in_use"], info["band_shape"], info["balanced_split"],
            info["n_ghost"], info["m_ghost"], int(np.sum(np.signbit(x) & (x == 0))), int(np.sum(np.signbit(y) & (y == 0)))), flush=True)
dist.barrier()
dist.destroy_process_group()
