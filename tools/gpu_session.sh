#!/bin/bash
# One gpurun call that brings back everything a round needs from the hardware (about 12-15 minutes of box time):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02'
#
# Everything lands in gpurun_out/<tag>_*; summarise with tools/ncu_summary.py and copy what is to be judged
# into profiles/.  Every step runs under its own timeout so that a hang costs minutes, not the box.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1

echo "== 1. GPU tests" | tee $out/${tag}_session.log
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $out/${tag}_session.log
tail -5 $out/${tag}_pytest_gpu.log | tee -a $out/${tag}_session.log

echo "== 2. smoke + bench (N = 1)" | tee -a $out/${tag}_session.log
timeout 120 python __graft_entry__.py --smoke >> $out/${tag}_session.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?" | tee -a $out/${tag}_session.log

echo "== 3. every kernel variant on the bench workload (CUDA events)" | tee -a $out/${tag}_session.log
for v in 1 2 3 4 5 6 7; do
  timeout 300 python tools/quick_bench.py --size 4096 --iters 100 --reps 3 --variant $v >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err
done
for f in 3 11; do  # compressed storage: every variant as well
  for v in 1 3 5 6 7; do
    timeout 300 python tools/quick_bench.py --size 4096 --iters 100 --reps 3 --variant $v --flags $f >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err
  done
done
timeout 300 python tools/quick_bench.py --kind random --size 4000000 --iters 50 --reps 3 >> $out/${tag}_variants.jsonl 2>> $out/${tag}_variants.err

echo "== 4. ncu launch list of the bench command" | tee -a $out/${tag}_session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --iters-per-step 10 --e2e-steps 0 --variants 0 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1

echo "== 5. ncu --set full of the hot kernels (chosen variants, and variant 1 for reference)" | tee -a $out/${tag}_session.log
# (launches of k_primal / k_dual in quick_bench: 70 while the variants are timed at creation, 40 warm-up, 20 timed,
#  32 with events between the kernels: skip past the first two groups)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_primal|k_dual' --launch-skip 114 -c 4 \
  -f -o $out/${tag}_hot python tools/quick_bench.py --size 4096 --iters 10 --reps 1 > $out/${tag}_ncu_hot.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_primal|k_dual' --launch-skip 44 -c 4 \
  -f -o $out/${tag}_hot_v1 python tools/quick_bench.py --size 4096 --iters 10 --reps 1 --variant 1 > $out/${tag}_ncu_hot_v1.log 2>&1

echo "== 6. tiny LPs: CUDA graphs vs the persistent CTA (printed by the test)" | tee -a $out/${tag}_session.log
timeout 300 python -m pytest tests/test_gpu_zz_opt_in_features.py -m gpu -q -s -k sc105_regression > $out/${tag}_tiny.log 2>&1
timeout 300 python bench.py --size 256 --steps 2 --warmup 3 --e2e-steps 0 --variants 0 --no-cpu-baseline --small-configs 2 > $out/${tag}_small_configs.json 2>> $out/${tag}_tiny.log
grep "SC105" $out/${tag}_tiny.log | tee -a $out/${tag}_session.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv >> $out/${tag}_session.log 2>&1
echo "== done" | tee -a $out/${tag}_session.log
