"""Where does an end-to-end call on N GPUs spend its time?  (scratch tool; run under torchrun)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
        tools/e2e_breakdown_dist.py [potts size]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import build_workload
from pysparselp_b200 import generators
from pysparselp_b200.ChambollePockPPD import make_solver, chambolle_pock_ppd

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lp, keep = build_workload("potts", size, pinned=True)
args = generators.lp_args(lp)


def now():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    if world > 1:
        dist.barrier()
    t0 = now()
    s = make_solver(*args)
    t1 = now()
    s.primal_step(keep_d=True); s.sync(); s.stats_step(False); st = s.read_stats(); s.dual_step(); s.sync()
    t2 = now()
    s.iterate(499); s.sync()
    t3 = now()
    x = s.get_x()
    t4 = now()
    st = s.read_stats_or_none()
    s.close()
    t5 = now()
    del x
    if rank == 0:
        print("rep %d N=%d: make_solver %.3f | stats iteration %.3f | 499 iterations %.3f | get_x %.3f | close %.3f | total %.3f s" % (
            rep, world, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0), flush=True)
for rep in range(3):
    if world > 1:
        dist.barrier()
    t0 = now()
    x, best = chambolle_pock_ppd(*args, nb_max_iter=500, nb_iter_plot=500)
    t1 = now()
    x = None
    if rank == 0:
        print("rep %d N=%d: chambolle_pock_ppd(500 iterations) %.3f s -> %.1f it/s" % (rep, world, t1 - t0, 500 / (t1 - t0)), flush=True)
if world > 1:
    dist.destroy_process_group()
