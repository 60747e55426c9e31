#!/bin/bash
# Debug aid: per-role wait/busy cycles of CTA 0 of the forward tensor-core kernel.
#   here (no GPU):   tools/role_profile.sh build    -> slak_b200/libslak_b200_prof.so  (-DSLAK_ROLE_PROFILE)
#   on the GPU box:  tools/role_profile.sh run      -> loads it through SLAK_B200_LIB (the product .so is untouched)
set -e
if [ "$1" = build ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DSLAK_ROLE_PROFILE \
    -Xcompiler -fPIC -shared -cudart static -o slak_b200/libslak_b200_prof.so slak_b200/csrc/*.cu -lcuda 2>&1 | grep -i "error" || true
  ls -la slak_b200/libslak_b200_prof.so
else
  SLAK_B200_LIB=$PWD/slak_b200/libslak_b200_prof.so python tools/role_profile.py 2>&1 | grep "warp\|stage" | sort -k1,1 -k3,3n | head -120
fi
