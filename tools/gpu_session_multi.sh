#!/bin/bash
# Multi-GPU companion of tools/gpu_session.sh (charged N x the box time — keep it short):
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_multi.sh r02 2'
#
# Parity worker on N GPUs, then bench.py with the three halo transports: peer-memory push / wait kernels (default),
# NCCL send/recv (flag 32), halo fused into k_primal / k_dual (flag 128 — opt-in until this run says otherwise; its
# waits give up after CPPPD_HALO_TIMEOUT_S instead of hanging the GPU).
tag=${1:-rXX}
n=${2:-2}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1 CPPPD_HALO_TIMEOUT_S=20
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }

echo "== parity worker on $n GPUs" | tee $out/${tag}_multi.log
run 600 29611 tests/dist_worker.py > $out/${tag}_dist_worker_n$n.log 2>&1
echo "exit $?" | tee -a $out/${tag}_multi.log
grep "DIST_WORKER_OK" $out/${tag}_dist_worker_n$n.log | tee -a $out/${tag}_multi.log

port=29621
for flags in 0 32 128; do
  echo "== bench --gpus $n --flags $flags" | tee -a $out/${tag}_multi.log
  run 600 $port bench.py --gpus $n --steps 10 --warmup 3 --e2e-steps 1 --flags $flags \
    > $out/${tag}_bench_n${n}_f$flags.json 2> $out/${tag}_bench_n${n}_f$flags.err
  echo "exit $?" | tee -a $out/${tag}_multi.log
  port=$((port + 1))
done
echo "== done" | tee -a $out/${tag}_multi.log
