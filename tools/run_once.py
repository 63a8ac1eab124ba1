"""run the bench workload a few times (for ncu captures): python tools/run_once.py [reps] [iters]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from b2 import pkg
w = bench.WORKLOAD
period, expected, flen = bench.make_period()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = torch.from_numpy(np.tile(period, reps).view(np.float32)).cuda()
rx = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], max_batch=len(period) * reps)
for _ in range(iters):
    rx.execute_device(x.data_ptr(), len(period) * reps)
    recs, pl = rx.poll()
print("frames", len(recs), "timing", rx.last_timing())
