#!/bin/bash
cd "$(dirname "$0")/.."
python tools/dbg_optim.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_merge_decom.py tests/test_masking_gpu.py tests/test_optim_gpu.py -m gpu -q 2>&1 | grep -v "Warning\|warn\|cosine_stepper\|^$" | tail -30
