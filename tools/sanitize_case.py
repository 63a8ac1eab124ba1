"""small end-to-end receive through the production kernels, for compute-sanitizer:
   compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_case.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refmc
from b2 import pkg
N, M, cp, taper, plen = 32, 512, 64, 16, 300
L = refmc.ref_lib()
tx = refmc.McTx(L, N, M, cp, taper)
x = tx.run((M + cp) * 24, plen, refmc.MOD_QAM64, refmc.FEC_NONE, refmc.FEC_NONE, max_frames=2, gain=1.0 / N)
tx.close()
orx = refmc.McRx(L, N, M, cp, taper); orx.execute(x); fo, po = orx.frames(); orx.close()
rx = pkg.MultichannelRx(N, M, cp, taper, device=0)
h = len(x) // 2 + 77
rx.execute(x[:h]); rx.execute(x[h:])
fg, pg = rx.poll()
rx.close()
assert len(fg) == len(fo) == 2 * N, (len(fg), len(fo))
assert np.array_equal(po, pg)
print("sanitize case ok:", len(fg), "frames")
# a conv-coded shape: Viterbi decode, worker pairs of the M = 256 synchroniser (one warp per worker)
N, M, cp, taper, plen = 8, 256, 32, 8, 200
tx = refmc.McTx(L, N, M, cp, taper)
x = tx.run((M + cp) * 60, plen, refmc.MOD_QAM16, refmc.FEC_CONV_V27, refmc.FEC_NONE, max_frames=3, gain=1.0 / N)
tx.close()
orx = refmc.McRx(L, N, M, cp, taper); orx.execute(x); fo, po = orx.frames(); orx.close()
rx = pkg.MultichannelRx(N, M, cp, taper, device=0)
h = len(x) // 3 + 5
rx.execute(x[:h]); rx.execute(x[h:])
fg, pg = rx.poll()
rx.close()
assert len(fg) == len(fo) == 3 * N, (len(fg), len(fo))
assert np.array_equal(po, pg) and int(fg["payload_valid"].min()) == 1
print("sanitize case 2 ok:", len(fg), "frames")
