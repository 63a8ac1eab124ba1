"""BASELINE configs[1] on the GPU: 8 channels, M = 64, QPSK, Hamming(12,8), device-resident input (python tools/c2_rate.py)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, refmc
from b2 import pkg
N, M, cp, taper, plen = 8, 64, 16, 4, 1200
L = refmc.ref_lib()
enc = (plen + 4) * 12 // 8
nd = 44
nsym = 3 + -(-288 // nd) + -(-(8 * enc // 2) // nd) + 1
flen = nsym * (M + cp)
tx = refmc.McTx(L, N, M, cp, taper)
x = tx.run(2 * flen, plen, refmc.MOD_QPSK, refmc.FEC_NONE, refmc.FEC_HAMMING128, seed=0xB2000000, gain=1.0 / N)
tx.close()
K = 2 * N
period = x[flen * K:2 * flen * K].copy()
reps = 256
n = len(period) * reps
d = torch.from_numpy(period.view(np.float32)).cuda().repeat(reps).contiguous()
rx = pkg.MultichannelRx(N, M, cp, taper, max_batch=n)
for _ in range(2):
    rx.execute_device(d.data_ptr(), n); recs, pl = rx.poll()
print("frames", len(recs), "valid", int(recs["payload_valid"].sum()))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    rx.execute_device(d.data_ptr(), n); rx.poll_view()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print("C2: %d samples/step, %.2f ms, %.1f Msamples/s" % (n, dt * 1e3, n / dt / 1e6), rx.last_timing())
