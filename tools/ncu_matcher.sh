#!/bin/sh
# `ncu --set full` of the matcher microbench kernels (cost + warp-per-image LSAP) on one 16384-image chunk at T = 100 and
# T = 10 (dev tool; run on the GPU box; text summaries only).
tag=${1:-r02}
for T in 10 100; do
  for k in "matcher_cost:cost" "lsap_kernel:lsap"; do
    re=$(echo "$k" | cut -d: -f1); name=$(echo "$k" | cut -d: -f2)
    timeout 200 ncu --set full --clock-control none -k "regex:$re" --launch-skip 2 --launch-count 1 -o /tmp/m_${name}_$T python tools/run_matcher.py $T > /dev/null 2>&1
    python tools/summarize_ncu.py /tmp/m_${name}_$T.ncu-rep > gpurun_out/${tag}_ncu_full_matcher_${name}_T$T.txt 2>&1
    rm -f /tmp/m_${name}_$T.ncu-rep
  done
done
ls -la gpurun_out/${tag}_ncu_full_matcher_*
