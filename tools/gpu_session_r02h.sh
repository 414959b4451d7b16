#!/bin/bash
# Round 2, session h (1 GPU): L2 window prefetch A/B on the banded random LP, banded tests
tag=r02h
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. banded GPU tests" | tee $log
timeout 600 python -m pytest tests/test_gpu_banded.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_gpu.log | tee -a $log
echo "== 2. prefetch on / off, windows 48 / 56 / 64 MB" | tee -a $log
for pf in 1 0; do for mb in 48 56 64; do
  CPPPD_BAND_PREFETCH=$pf CPPPD_BAND_WINDOW_MB=$mb timeout 300 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 --flags 1024 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
done; done
echo "== done" | tee -a $log
