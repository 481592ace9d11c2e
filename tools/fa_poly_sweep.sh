#!/bin/sh
# A/B of the polynomial-exp2 fraction of the fused attention forward (OWL_FA_POLY = exponentials per 8 on the FMA pipe):
# parity tests + timing at the two bench shapes for every setting.  usage: sh tools/fa_poly_sweep.sh [outfile]
out=${1:-gpurun_out/fa_poly_sweep.txt}
mkdir -p gpurun_out
: > $out
for p in 0 2 3 4; do
  echo "== OWL_FA_POLY=$p" >> $out
  OWL_FA_POLY=$p timeout 120 python -m pytest tests/test_flash_attn_gpu.py -q -x -k "fwd or peaky" 2>&1 | tail -1 >> $out
  OWL_FA_POLY=$p timeout 60 python tools/time_flash.py 16 577 12 >> $out 2>&1
  OWL_FA_POLY=$p timeout 60 python tools/time_flash.py 4 3601 16 >> $out 2>&1
done
cat $out
