#!/bin/bash
# Round 2, session w (1 GPU): cluster kernel v2 (precomputed shared::cluster addresses, state in registers)
tag=r02x
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_zz_opt_in_features.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
for mode in 0 1; do
  echo "-- CPPPD_CLUSTER_MODE=$mode"
  CPPPD_CLUSTER_MODE=$mode timeout 300 python tools/small_bench.py 5000 2>&1 | tail -5 | tee -a $out/${tag}_small_mode$mode.jsonl
done
echo "-- forced"
CPPPD_FORCE_CLUSTER=1 timeout 300 python tools/small_bench.py 5000 2>&1 | tail -7 | head -2 | cut -c1-400
