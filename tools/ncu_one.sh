#!/bin/bash
# On the GPU box: ncu --set full on ONE kernel of a microbenchmark run, exported as raw + source CSV pages.
# usage: tools/ncu_one.sh <tag> <kernel-regex> <launch-skip> -- <command...>
set -u
TAG=$1; KRE=$2; SKIP=$3; shift 4
OUT=gpurun_out
REP=/tmp/$TAG.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$KRE" --launch-skip $SKIP --launch-count 1 \
  -f -o /tmp/$TAG "$@" > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
ncu -i $REP --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_raw.csv.gz
ncu -i $REP --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $OUT/${TAG}_sass.csv.gz
ls -la $OUT | tail -5
