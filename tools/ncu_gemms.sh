#!/bin/sh
# `ncu --set full` capture of one launch of each encoder-layer GEMM inside an eager step (dev tool; run on the GPU box).
# Template arguments are part of the match, hence --kernel-name-base demangled.  Only text summaries are kept.
tag=${1:-r02}
for k in "EpiF32, .int.2, .int.1>:fc2:14" "EpiF16<.int.1>, .int.1, .int.1>:fc1:14" "192, .bool.0, .bool.0, owl::EpiF16<.int.0>, .int.1, .int.1>:qkv:14" "192, .bool.0, .bool.0, owl::EpiF32, .int.1, .int.1>:out_proj:14"; do
  re=$(echo "$k" | cut -d: -f1-1); name=$(echo "$k" | awk -F: '{print $(NF-1)}'); skip=$(echo "$k" | awk -F: '{print $NF}')
  re=$(echo "$k" | sed "s/:$name:$skip\$//")
  timeout 200 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$re" --launch-skip $skip --launch-count 1 -o /tmp/full_$name python tools/prof_step.py 16 2 > /tmp/ncu_$name.log 2>&1
  python tools/summarize_ncu.py /tmp/full_$name.ncu-rep > gpurun_out/${tag}_ncu_full_$name.txt 2>&1 || tail -5 /tmp/ncu_$name.log
  rm -f /tmp/full_$name.ncu-rep
done
ls -la gpurun_out/${tag}_ncu_full_*
