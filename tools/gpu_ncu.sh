#!/bin/bash
# round-2 ncu evidence: (1) launch list of a few eager steps, (2) DRAM bytes + tensor-pipe activity of every kernel of the step,
# (3) --set full captures of the dominant kernels (raw pages exported here: a whole-step .ncu-rep exceeds the 64 MiB limit)
set -u
O=gpurun_out
mkdir -p $O
CMD="python bench.py --graph off --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-kernel-table"
timeout 300 $CMD > $O/ncu_plain.json 2> $O/ncu_plain.err || { echo "plain run failed"; tail -5 $O/ncu_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 800 --csv --log-file $O/r02_launches_r224_tf32.csv $CMD > $O/ncu_l.log 2>&1; echo "launch list rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 1200 ncu --metrics $M --clock-control none -s 1000 -c 480 --csv --log-file $O/r02_step_metrics_r224_tf32.csv $CMD > $O/ncu_m.log 2>&1; echo "step metrics rc=$?"
K='regex:conv_halo_tc_k|conv_fprop_tc_k|conv_wgrad_tc_k|conv_wgrad_halo_k|conv_wgrad_strip_k|conv_pers_tc_k|bn_bwd_apply_persist_k|col_sums_atomic_k|bn_bwd_fused_k|bn_apply_persist_k|pool_bn_|bn_relu_maxpool_k3s2_k'
timeout 1500 ncu --set full --clock-control none -k "$K" -s 125 -c 125 -o $O/r02_full $CMD > $O/ncu_f.log 2>&1; echo "full rc=$?"
ncu -i $O/r02_full.ncu-rep --page raw --csv > $O/r02_ncu_full_raw.csv 2>/dev/null; ls -la $O/r02_full.ncu-rep $O/r02_ncu_full_raw.csv
rm -f $O/r02_full.ncu-rep
tail -2 $O/ncu_l.log $O/ncu_m.log $O/ncu_f.log
