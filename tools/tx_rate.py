"""multichanneltx throughput on the GPU: every channel re-armed at each frame boundary, samples left on the
device (python tools/tx_rate.py [N] [M] [cp] [taper] [frames])"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from b2 import pkg
N, M, cp, taper, frames = [int(a) for a in (sys.argv[1:6] + [256, 512, 64, 16, 8][len(sys.argv) - 1:])]
K = 2 * N
tx = pkg.MultichannelTx(N, M, cp, taper)
rng = np.random.default_rng(1)
payload = rng.integers(0, 256, 1200, dtype=np.uint8)
hdr = np.arange(8, dtype=np.uint8)
MOD_QAM64, FEC_NONE = 29, 1
nsym = 3 + 1 + 5 + 1
calls_per_frame = nsym * (M + cp)
out = torch.empty(calls_per_frame * K * 2, dtype=torch.float32, device="cuda")
tot = [0.0, 0.0, 0.0, 0.0]
t0 = None
for f in range(frames + 2):
    if f == 2:
        torch.cuda.synchronize(); t0 = time.perf_counter(); tot = [0.0] * 4
    for c in range(N):
        if tx.is_ready(c):
            tx.update(c, hdr, payload, MOD_QAM64, FEC_NONE, FEC_NONE)
    tx.generate_device(out.data_ptr(), calls_per_frame)
    tm = tx.last_timing()
    tot = [a + b for a, b in zip(tot, tm)]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n = frames * calls_per_frame * K
print("N=%d M=%d: %.1f Msamples/s wall (%d samples in %.1f ms); kernel ms per frame period %s" % (N, M, n / dt / 1e6, n, 1e3 * dt, [round(x / frames, 3) for x in tot]))
