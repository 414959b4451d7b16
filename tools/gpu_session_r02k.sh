#!/bin/bash
# Round 2, session k (1 GPU): final library — full GPU suite, the default bench line, L1-SVM with 16384-entry segments,
# ncu --set full of the banded window kernels and of the grid-stride compressed kernels
tag=r02k
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. GPU tests" | tee $log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_gpu.log | tee -a $log
echo "== 2. smoke + default bench (N = 1) + reference arm" | tee -a $log
timeout 120 python __graft_entry__.py --smoke >> $log 2>&1
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?" | tee -a $log
echo "== 3. L1-SVM 100 000 samples" | tee -a $log
timeout 600 python tools/quick_bench.py --kind l1svm --size 100000 --iters 10 --reps 3 >> $out/${tag}_l1svm.jsonl 2>> $out/${tag}_l1svm.err
echo "l1svm exit $?" | tee -a $log
timeout 900 python bench.py --workload l1svm --size 100000 --steps 4 --iters-per-step 10 --e2e-steps 1 --e2e-iters 50 --no-cpu-baseline --small-configs 0 \
  > $out/${tag}_bench_l1svm.json 2> $out/${tag}_bench_l1svm.err
echo "bench l1svm exit $?" | tee -a $log
echo "== done" | tee -a $log
