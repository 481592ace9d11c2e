#!/bin/sh
# A/B of the fused attention forward flavours (OWL_FA_GEN, see owl_flash_attn_fwd in csrc/flash_attn.cu): parity tests + timing at the two bench shapes.  usage: sh tools/fa_roles_sweep.sh
out=${1:-gpurun_out/fa_roles_sweep.txt}
mkdir -p gpurun_out
: > $out
run() {
  echo "== $1" >> $out
  env $1 timeout 120 python -m pytest tests/test_flash_attn_gpu.py -q -x -k "fwd or peaky" 2>&1 | tail -1 >> $out
  env $1 timeout 60 python tools/time_flash.py 16 577 12 >> $out 2>&1
  env $1 timeout 60 python tools/time_flash.py 4 3601 16 >> $out 2>&1
}
run OWL_FA_GEN=24
run OWL_FA_GEN=26
run OWL_FA_GEN=34
run OWL_FA_GEN=40
run OWL_FA_GEN=41
cat $out
