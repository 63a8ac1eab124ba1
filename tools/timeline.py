"""stage timeline of one device-resident call of the bench workload (B2_DUMP_TIMELINE output on stderr):
python tools/timeline.py [reps]"""
import os, sys
os.environ["B2_DUMP_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, time
import bench
from b2 import pkg
w = bench.WORKLOAD
period, expected, flen = bench.make_period()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 91
x = torch.from_numpy(period.view(np.float32)).cuda().repeat(reps).contiguous()
n = len(period) * reps
rx = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], max_batch=n)
for i in range(4):
    if i == 3:
        sys.stderr.write("---- timed call\n")
    t0 = time.perf_counter()
    rx.execute_device(x.data_ptr(), n)
    t1 = time.perf_counter()
    recs, pl = rx.poll_view()
    t2 = time.perf_counter()
sys.stderr.write("host: execute %.3f ms, poll_view %.3f ms, frames %d\n" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), len(recs)))
