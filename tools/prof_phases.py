"""per-phase cycle profile of the register-resident synchroniser (CTA 0, one launch of the bench workload):
   make -C liquid-usrp_b200 prof && python tools/prof_phases.py   (swaps the profile build in for this process only)"""
import sys, os, ctypes as C, shutil
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
shutil.copy("/root/repo/liquid-usrp_b200/libb200ofdm_prof.so", "/root/repo/liquid-usrp_b200/libb200ofdm.so")
import numpy as np, torch
import bench
from b2 import pkg
w = bench.WORKLOAD
period, expected, flen = bench.make_period()
reps = 8
x = torch.from_numpy(np.tile(period, reps).view(np.float32)).cuda()
rx = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], max_batch=len(period) * reps)
L = pkg.lib()
for _ in range(3):
    rx.execute_device(x.data_ptr(), len(period) * reps); rx.poll()
out = (C.c_ulonglong * 16)()
L.b2_debug_sync8_prof(out, 1)
rx.execute_device(x.data_ptr(), len(period) * reps); rx.poll()
L.b2_debug_sync8_prof(out, 0)
names = ["top wait", "consume+pass1", "fft rest", "eq+pilots", "fit: barrier wait", "derot+demap(+next consume)", "emit (record, copy, reset)", "-", "fit: atan2", "fit: unwrap", "fit: sums", "fit: p0/p1/nco", "preamble (all)", "header: demap+evm+bits", "header decode / payload tail", "launch set-up"]
tot = sum(out[:16])
print("timing", rx.last_timing())
for i, n in enumerate(names):
    print("%-14s %10d cycles %5.1f%%" % (n, out[i], 100.0 * out[i] / tot))
print("total cycles", tot, "-> ms at 1.9GHz", tot / 1.9e6)
