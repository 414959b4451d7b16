#!/bin/bash
# Round 2, session t (N GPUs): parity worker, end-to-end breakdown (sliced upload + all-gather), bench line
tag=r02t
n=${1:-2}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1 CPPPD_HALO_TIMEOUT_S=20
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 600 29611 tests/dist_worker.py > $out/${tag}_dist_worker_n$n.log 2>&1
echo "dist worker exit $?"; grep "DIST_WORKER_OK\|digest ok" $out/${tag}_dist_worker_n$n.log
CPPPD_SETUP_TIMING=1 run 300 29655 tools/e2e_breakdown_dist.py 4096 > $out/${tag}_e2e_breakdown_n$n.log 2>&1
echo "breakdown exit $?"; grep "rep \|rank 0\]" $out/${tag}_e2e_breakdown_n$n.log | tail -34
CPPPD_FULL_UPLOAD=1 run 300 29656 tools/e2e_breakdown_dist.py 4096 2>&1 | grep "rep " | tail -4
run 600 29621 bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_potts_n$n.json 2> $out/${tag}_bench_potts_n$n.err
echo "bench exit $?"
python - <<P
import json
d=json.loads(open("$out/${tag}_bench_potts_n$n.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["call_seconds"], d["parity"]["status"], d["gpu_launches"])
P
