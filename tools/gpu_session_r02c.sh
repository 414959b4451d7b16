#!/bin/bash
# Round 2, session c (1 GPU): where do the L2 sectors of the banded kernels come from (ncu --set full), L2 fetch granularity
tag=r02c
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. banded GPU tests" | tee $log
timeout 600 python -m pytest tests/test_gpu_banded.py -m gpu -x -q > $out/${tag}_pytest_banded.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_banded.log | tee -a $log
echo "== 2. ncu --set full: one middle window of each banded kernel + the last ones" | tee -a $log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_primal_band|k_dual_band' --launch-skip 113 -c 8 \
  -f -o $out/${tag}_band python tools/quick_bench.py --kind random --size 20000000 --iters 4 --reps 1 --flags 1024 > $out/${tag}_ncu_band.log 2>&1
echo "ncu exit $?" | tee -a $log
echo "== 3. SELL kernels on the random LP with 32-byte L2 fetch granularity" | tee -a $log
CPPPD_L2_FETCH_GRANULARITY=32 timeout 300 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 --flags 2048 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
timeout 300 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 --flags 2048 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
CPPPD_L2_FETCH_GRANULARITY=32 timeout 300 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 --flags 1024 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
echo "== done" | tee -a $log
