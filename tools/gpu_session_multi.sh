#!/bin/bash
# Multi-GPU session (charged N x the box time — keep it short):
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_multi.sh r02 2'
#
# Parity worker on N GPUs, then bench.py on the Potts LP (peer-memory push / wait kernels, and the halo fused into
# k_primal / k_dual, flag 128) and on the random LP (banded operands + balanced split).  Every bench line carries the
# parity verdict of its configuration (digest of x, y against the C port's).
tag=${1:-rXX}
n=${2:-2}
what=${3:-all}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1 CPPPD_HALO_TIMEOUT_S=20 CPPPD_SETUP_TIMING=1
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }

if [ "$what" = all ] || [ "$what" = parity ]; then
  echo "== parity worker on $n GPUs" | tee $out/${tag}_multi_n$n.log
  run 900 29611 tests/dist_worker.py > $out/${tag}_dist_worker_n$n.log 2>&1
  echo "exit $?" | tee -a $out/${tag}_multi_n$n.log
  grep "DIST_WORKER_OK\|digest ok" $out/${tag}_dist_worker_n$n.log | tee -a $out/${tag}_multi_n$n.log
fi
port=29621
if [ "$what" = all ] || [ "$what" = potts ]; then
  for flags in 0 ${POTTS_EXTRA_FLAGS:-}; do
    echo "== bench potts --gpus $n --flags $flags" | tee -a $out/${tag}_multi_n$n.log
    run 600 $port bench.py --gpus $n --steps 10 --warmup 3 --e2e-steps 1 --flags $flags \
      > $out/${tag}_bench_potts_n${n}_f$flags.json 2> $out/${tag}_bench_potts_n${n}_f$flags.err
    echo "exit $?" | tee -a $out/${tag}_multi_n$n.log
    port=$((port + 1))
  done
fi
if [ "$what" = all ] || [ "$what" = random ]; then
  echo "== bench random --gpus $n" | tee -a $out/${tag}_multi_n$n.log
  run 900 $port bench.py --workload random --gpus $n --steps 6 --warmup 3 --e2e-steps 1 --e2e-iters 200 --iters-per-step 20 \
    > $out/${tag}_bench_random_n${n}.json 2> $out/${tag}_bench_random_n${n}.err
  echo "exit $?" | tee -a $out/${tag}_multi_n$n.log
fi
if [ "$what" = all ] || [ "$what" = ngpus ]; then
  echo "== one process, $n GPUs (no torchrun): chambolle_pock_ppd(..., n_gpus=$n)" | tee -a $out/${tag}_multi_n$n.log
  timeout 600 python tools/n_gpus_check.py $n > $out/${tag}_n_gpus_check_n$n.log 2>&1
  echo "exit $?" | tee -a $out/${tag}_multi_n$n.log
  grep "from one process\|regression\|N_GPUS_CHECK_OK" $out/${tag}_n_gpus_check_n$n.log | tee -a $out/${tag}_multi_n$n.log
fi
echo "== done" | tee -a $out/${tag}_multi_n$n.log
