#!/bin/bash
# Round 2, session d (1 GPU): packed banded layout — tests, window sweep, request-rate metrics
tag=r02g
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
log=$out/${tag}_session.log
echo "== 1. GPU tests" | tee $log
timeout 900 python -m pytest tests/test_gpu_banded.py -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $log
tail -3 $out/${tag}_pytest_gpu.log | tee -a $log
echo "== 2. random LP 20M x 40M: window sweep (banded forced, flag 1024)" | tee -a $log
for mb in 48 56 64; do
  CPPPD_BAND_WINDOW_MB=$mb timeout 300 python tools/quick_bench.py --kind random --size 20000000 --iters 20 --reps 3 --flags 1024 >> $out/${tag}_random.jsonl 2>> $out/${tag}_random.err
done
echo "== 3. ncu of the banded kernels (selected metrics)" | tee -a $log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex.sum,lts__t_sectors_srcunit_ltcfabric.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct \
  --clock-control none -k regex:'k_primal_band|k_dual_band' --launch-skip 900 -c 11 --csv --log-file $out/${tag}_random_ncu.csv \
  python tools/quick_bench.py --kind random --size 20000000 --iters 4 --reps 1 --flags 1024 > $out/${tag}_random_ncu.log 2>&1
echo "ncu exit $?" | tee -a $log
echo "== done" | tee -a $log
