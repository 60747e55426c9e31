#!/bin/bash
# On the GPU box: one `ncu --set full` pass over every kernel of this repo in one Block per stage (fwd+bwd),
# then export what profiles/ keeps (the .ncu-rep of ~60 kernels is too large to bring back whole):
#   <tag>_raw.csv.gz          raw page (all metrics, one row per launch)
#   <tag>_sass_<kernel>.csv.gz  SASS page with stall samples for the kernels named below (first launch of each)
# usage: tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r02_block}
OUT=gpurun_out
REP=/tmp/$TAG.ncu-rep
timeout 800 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"^(lk|bn3|residual|gelu_bwd|wgrad3|mlp_gemm|colsum|cast_transpose)" -f -o /tmp/$TAG python tools/ncu_block.py 2 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
ncu -i $REP --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_raw.csv.gz
for K in lk3_fwd_tc_kernel lk_dgrad_tc_kernel lk3_wgrad_tc_kernel bn3_sum_ln_bwd_kernel bn3_sum_ln_fwd_kernel mlp_gemm_nt_kernel mlp_gemm_tn_splitk_kernel; do
  # launches of one kernel are ordered by stage: 1 = 56x56 (T=64 class), 3 = 14x14 (T=16 class)
  for SKIP in 0 2; do
    S=$SKIP
    if [ $K = lk_dgrad_tc_kernel ]; then S=$((2 * SKIP + 1)); fi   # two launches per Block: the second is the large one
    ncu -i $REP --page source --csv --print-source cuda,sass -k regex:"^$K" --launch-skip $S --launch-count 1 2>/dev/null \
      | gzip -9 > $OUT/${TAG}_sass_${K}_s$((SKIP / 2 * 2 + 1)).csv.gz
  done
done
ls -la $OUT | tail -20
du -sh $OUT
