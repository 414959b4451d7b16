#!/bin/bash
# Round 2, session l (1 GPU): ncu --set full of the banded window kernels (random LP), the grid-stride kernels (Potts,
# generic and compressed storage) and the long-row kernels (L1-SVM).  (No --import-source: gpurun brings back <= 64 MiB.)
tag=r02l
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== ncu captures" | tee $log
timeout 900 ncu --set full --clock-control none -k regex:'k_primal_band|k_dual_band' --launch-skip 560 -c 9 \
  -f -o $out/${tag}_band python tools/quick_bench.py --kind random --size 20000000 --iters 4 --reps 1 > $out/${tag}_ncu_band.log 2>&1
echo "ncu band exit $?" | tee -a $log
timeout 600 ncu --set full --clock-control none -k regex:'k_primal|k_dual' --launch-skip 44 -c 2 \
  -f -o $out/${tag}_stride python tools/quick_bench.py --size 4096 --iters 10 --reps 1 --variant 8 > $out/${tag}_ncu_stride.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_primal|k_dual' --launch-skip 44 -c 2 \
  -f -o $out/${tag}_stride_dict python tools/quick_bench.py --size 4096 --iters 10 --reps 1 --variant 8 --flags 11 > $out/${tag}_ncu_stride_dict.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_long_partial|k_long_finish' --launch-skip 20 -c 2 \
  -f -o $out/${tag}_long python tools/quick_bench.py --kind l1svm --size 50000 --iters 4 --reps 1 > $out/${tag}_ncu_long.log 2>&1
echo "ncu exit $?" | tee -a $log
ls -la $out | tee -a $log
echo "== done" | tee -a $log
