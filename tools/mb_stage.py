"""One stage geometry through tools/microbench.py: python tools/mb_stage.py <fwd_tc|bwd_tc|simt|all> N C H KL"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.microbench as mb  # noqa: E402

N, C, H, KL = [int(a) for a in sys.argv[2:6]]
mb.stage(N, C, H, KL, sys.argv[1])
