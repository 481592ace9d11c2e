#!/bin/sh
# Per-kernel ncu evidence of one eager train step at batch 16 (dev tool; run on the GPU box).  Writes only small text files
# under gpurun_out/ (the .ncu-rep files are summarised on the box and deleted: gpurun returns at most 64 MiB).
#   gpurun_out/<tag>_launches.csv         every launch of the second step with its device time (shares, cold-cache)
#   gpurun_out/<tag>_ncu_step.txt         judged metrics of every launch of the second step (explicit metric list)
#   gpurun_out/<tag>_ncu_full_<k>.txt     `--set full` capture of one launch of selected kernels
tag=${1:-r02}
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/prof_step.py 16 2 > /dev/null 2>&1
timeout 400 ncu --metrics $M --clock-control none --launch-skip 175 --launch-count 190 -o /tmp/step_metrics python tools/prof_step.py 16 2 > /dev/null 2>&1
python tools/summarize_ncu.py /tmp/step_metrics.ncu-rep > gpurun_out/${tag}_ncu_step.txt 2>&1
rm -f /tmp/step_metrics.ncu-rep
# kernel regex : name : launches to skip (12-per-step kernels: second step's third layer; once-per-step kernels: second step)
for k in "EpiF32.*2, .int.1>:fc2:14" "EpiF16<.int.1>:fc1:14" "flash_attn:flash:14" "post_fuse_kernel:post_fuse:1" "box_tail_kernel:box_tail:1" "matcher_cost:matcher_cost:1" "u8_patches:u8_patches:1" "rownorm_kernel:rownorm:2"; do
  re=$(echo "$k" | cut -d: -f1); name=$(echo "$k" | cut -d: -f2); skip=$(echo "$k" | cut -d: -f3)
  timeout 200 ncu --set full --clock-control none -k "regex:$re" --launch-skip $skip --launch-count 1 -o /tmp/full_$name python tools/prof_step.py 16 2 > /dev/null 2>&1
  python tools/summarize_ncu.py /tmp/full_$name.ncu-rep > gpurun_out/${tag}_ncu_full_$name.txt 2>&1
  rm -f /tmp/full_$name.ncu-rep
done
ls -la gpurun_out/ | head -40
