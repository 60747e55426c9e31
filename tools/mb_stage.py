import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.microbench as mb
N, C, H, KL = [int(a) for a in sys.argv[2:6]]
mb.stage(N, C, H, KL, sys.argv[1])
