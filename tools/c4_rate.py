"""BASELINE configs[3] on the GPU: single-link frames (M = 512, cp 64, 256-QAM, no FEC, 1200-byte payloads), a batch of
4096 of them: ofdmflexframegen -> msresamp_crcf(1.07) -> msresamp_crcf(1/1.07) -> ofdmflexframesync.  The frames of a
batch are independent (idle gaps between them), so the synchroniser takes them as 4096 streams (b2_ofdmsync, streams > 1)
instead of one serial chain.   B2_MAX_PAYLOAD=2048 python tools/c4_rate.py [nframes]"""
import ctypes as C
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("B2_MAX_PAYLOAD", "2048")
import numpy as np, torch
from b2 import pkg
M, cp, taper, plen = 512, 64, 16, 1200
nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = M + cp
L = pkg.lib()
rng = np.random.default_rng(3)
# a handful of distinct frames, tiled (the generator itself is measured in tests/test_gpu_tx.py)
g = pkg.OfdmGen(M, cp, taper)
kinds, sent = [], []
for f in range(8):
    header = rng.integers(0, 256, 8, dtype=np.uint8)
    payload = rng.integers(0, 256, plen, dtype=np.uint8)
    nsym = g.assemble(header, payload, 6, 1, 1, 31)          # CRC-32, fec none/none, 256-QAM
    out, last = g.write(nsym)
    kinds.append(np.concatenate([out, np.zeros(3 * W, np.complex64)]))
    sent.append(payload)
g.close()
seg = len(kinds[0])
x = np.concatenate([kinds[f % 8] for f in range(nframes)])
n = len(x)
d_x = torch.from_numpy(x.view(np.float32)).cuda()
d_y = torch.empty(2 * (int(n * 1.08) + 1024), dtype=torch.float32, device="cuda")
d_z = torch.empty(2 * (n + 4096), dtype=torch.float32, device="cuda")
up, down = pkg.MsResamp(np.float32(1.07)), pkg.MsResamp(np.float32(1.0 / 1.07))


def resample(h, src, nsrc, dst):
    ny = C.c_size_t(0)
    rc = L.b2_msresamp_execute_device(h.h, C.c_void_p(src.data_ptr()), C.c_size_t(nsrc), C.c_void_p(dst.data_ptr()), C.c_size_t(dst.numel() // 2), C.byref(ny))
    assert rc == 0
    return ny.value


rx = pkg.OfdmSync(M, cp, taper, streams=nframes, max_batch=seg)
best = {}
for it in range(4):
    up.reset(); down.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ny = resample(up, d_x, n, d_y)
    nz = resample(down, d_y, ny, d_z)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    # the two resamplers delay the waveform by a few samples; frames stay inside their own segment (3 idle symbols)
    rx.execute_device(d_z.data_ptr(), seg, seg)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    fr, pl = rx.poll()
    t3 = time.perf_counter()
    best = {"resample_ms": 1e3 * (t1 - t0), "sync_decode_ms": 1e3 * (t2 - t1), "poll_ms": 1e3 * (t3 - t2)}
ok = int(fr["payload_valid"].sum())
c = int(fr["channel"][0]); o = int(fr["payload_offset"][0])
assert np.array_equal(pl[o:o + plen], sent[c % 8])
tot = best["resample_ms"] + best["sync_decode_ms"]
print("C4: %d frames, %d generated samples; resample x2 %.2f ms, sync+decode %.2f ms (%d frames valid) -> %.1f Msamples/s; timing %s"
      % (nframes, n, best["resample_ms"], best["sync_decode_ms"], ok, n / tot / 1e3, rx.last_timing()))
