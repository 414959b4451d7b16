#!/bin/bash
# Round 2, first hardware session (1 GPU): what the design of the random-LP kernels needs to know, the
# "before" picture of configs[3], and fresh ncu evidence for the kernels the bench really runs.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_session_r02a.sh'
tag=r02a
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 0. box" | tee $log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv >> $log 2>&1
nproc >> $log; free -g | head -2 >> $log

echo "== 0b. GPU tests" | tee -a $log
timeout 600 python -m pytest tests -m gpu -x -q --durations=10 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_gpu.log | tee -a $log

echo "== 1. gather probe (L2 window sweep)" | tee -a $log
timeout 300 tools/probe/gather_probe > $out/${tag}_gather_probe.jsonl 2>&1
echo "probe exit $?" | tee -a $log

echo "== 2. random LP 20M x 40M: autotuned variants, CUDA events" | tee -a $log
timeout 600 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
echo "random exit $?" | tee -a $log
echo "== 3. random LP: L2 hit rate + DRAM bytes of variant 1 (ncu, selected metrics)" | tee -a $log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum \
  --clock-control none -k regex:'k_primal|k_dual' --launch-skip 40 -c 2 --csv --log-file $out/${tag}_random_ncu.csv \
  python tools/quick_bench.py --kind random --size 20000000 --iters 4 --reps 1 --variant 1 > $out/${tag}_random_ncu.log 2>&1
echo "random ncu exit $?" | tee -a $log

echo "== 4. ncu launch list of the bench command (Potts 4096^2)" | tee -a $log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --iters-per-step 10 --e2e-steps 0 --variants 0 --no-cpu-baseline --small-configs 0 > $out/${tag}_ncu_bench.log 2>&1
echo "launch list exit $?" | tee -a $log

echo "== 5. ncu --set full: autotuned generic kernels, and the dictionary kernels (flags 11)" | tee -a $log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_primal|k_dual' --launch-skip 114 -c 4 \
  -f -o $out/${tag}_hot python tools/quick_bench.py --size 4096 --iters 10 --reps 1 > $out/${tag}_ncu_hot.log 2>&1
echo "ncu hot exit $?" | tee -a $log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_primal|k_dual' --launch-skip 114 -c 4 \
  -f -o $out/${tag}_hot_dict python tools/quick_bench.py --size 4096 --iters 10 --reps 1 --flags 11 > $out/${tag}_ncu_hot_dict.log 2>&1
echo "ncu dict exit $?" | tee -a $log

echo "== 6. tiny LPs: CUDA graphs vs the persistent CTA" | tee -a $log
timeout 300 python bench.py --size 256 --steps 2 --warmup 3 --e2e-steps 0 --variants 0 --no-cpu-baseline --small-configs 2 > $out/${tag}_small_configs.json 2>> $out/${tag}_tiny.log
echo "tiny exit $?" | tee -a $log
echo "== done" | tee -a $log
