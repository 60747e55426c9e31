#!/bin/bash
# On the GPU box: one `ncu --set full` pass over every kernel of this repo in one Block per stage (fwd+bwd),
# then export what profiles/ keeps (the .ncu-rep of ~70 kernels is too large to bring back whole):
#   <tag>_raw.csv.gz            raw page (all metrics, one row per launch)
#   <tag>_sass_<kernel>_<i>.csv.gz  SASS page with stall samples of the i-th launch of the kernels named below
# usage: tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r02_block}
OUT=gpurun_out
REP=/tmp/$TAG.ncu-rep
timeout 1200 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"^(lk|bn3|residual|wgrad3|mlp_gemm|colsum|cast_transpose|dense_|ln_fwd2|ln_bwd2|res_fwd2|res_bwd2)" -f -o /tmp/$TAG python tools/ncu_block.py 2 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
ncu -i $REP --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_raw.csv.gz
# launches of one kernel are in execution order: Block 1 (56x56) forward, backward, Block 2, ...; dense_kernel: 0 = forward 14x14,
# 1 = data gradient 14x14, 2 / 3 = the same at 7x7
for SPEC in lk3_fwd_tc_kernel:0 lk_dgrad_tc_kernel:1 lk3_wgrad_tc_kernel:0 dense_kernel:0 dense_kernel:1 dense_wgrad_kernel:0 ln_bwd2_kernel:0 ln_bwd2_kernel:2 ln_fwd2_kernel:0 mlp_gemm_nt_kernel:0 mlp_gemm_nt_kernel:4; do
  K=${SPEC%%:*}; S=${SPEC##*:}
  ncu -i $REP --page source --csv --print-source cuda,sass -k regex:"^$K" --launch-skip $S --launch-count 1 2>/dev/null \
    | gzip -9 > $OUT/${TAG}_sass_${K}_$S.csv.gz
done
ls -la $OUT | tail -20
du -sh $OUT
