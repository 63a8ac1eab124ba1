"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
reference's L2 sources over the oracle).  CPU: the oracle still reproduces them.  GPU: the CUDA
path reproduces them through the C ABI without the oracle in the loop."""
import os

import numpy as np
import pytest

import orc
from refmc import McRx, McTx, ref_lib, CRC_32, FEC_NONE, FEC_HAMMING128, FEC_CONV_V27, MOD_QPSK, MOD_QAM16

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 0xB2000000
EXACT = ("channel", "header_valid", "payload_valid", "payload_len", "header", "mod_scheme", "mod_bps",
         "check", "fec0", "fec1", "detect_index", "complete_index", "payload_offset")


def same_frames(a, b):
    assert len(a) == len(b)
    for k in EXACT:
        assert np.array_equal(a[k], b[k]), k


def test_oracle_reproduces_packetizer_golden():
    g = np.load(os.path.join(G, "packetizer.npz"))
    for key in g.files:
        if key == "msg":
            continue
        f0, f1 = (int(v) for v in key[1:].split("_"))
        assert np.array_equal(orc.packetizer_encode(g["msg"], CRC_32, f0, f1), g[key]), key
        d, ok = orc.packetizer_decode(g[key], len(g["msg"]), CRC_32, f0, f1)
        assert ok and np.array_equal(d, g["msg"])


def test_oracle_reproduces_loopback_golden():
    g = np.load(os.path.join(G, "mc_n2_m64.npz"))
    tx = McTx(ref_lib(), 2, 64, 16, 4)
    x = tx.run(80 * 24, 60, MOD_QPSK, FEC_NONE, FEC_HAMMING128, seed=SEED, max_frames=1, gain=0.5)
    tx.close()
    assert np.array_equal(x, g["x"])
    rx = McRx(ref_lib(), 2, 64, 16, 4)
    rx.execute(g["x"])
    fr, pl = rx.frames()
    rx.close()
    same_frames(fr, g["frames"])
    assert np.array_equal(pl, g["payloads"])


@pytest.mark.gpu
def test_cuda_receiver_reproduces_golden():
    from b2 import pkg
    g = np.load(os.path.join(G, "mc_n2_m64.npz"))
    rx = pkg.MultichannelRx(2, 64, 16, 4)
    rx.tap_symbols(True, 4096)
    rx.execute(g["x"])
    fr, pl = rx.poll()
    ch, idx, X = rx.read_symbols()
    rx.close()
    same_frames(fr, g["frames"])
    assert np.array_equal(pl, g["payloads"])
    sel = ch == 0
    assert np.array_equal(idx[sel], g["sym_index"])
    err = np.abs(X[sel] - g["sym_X"]).max(axis=1) / np.abs(g["sym_X"]).max(axis=1)
    assert err.max() < 1e-5


@pytest.mark.gpu
def test_cuda_loopback_reproduces_golden_records():
    """CUDA transmitter -> CUDA receiver, compared with the records the reference produced"""
    from b2 import pkg
    g = np.load(os.path.join(G, "mc_n4_m256.npz"))
    N, M, cp, taper, plen = 4, 256, 32, 8, 200
    L = ref_lib()
    tx = pkg.MultichannelTx(N, M, cp, taper)
    pid = [0] * N
    out = []
    ncalls = int(g["ncalls"])
    done = 0
    while done < ncalls:                       # the src/multichannel_tx.cc loop
        for c in range(N):
            if pid[c] < 2 and tx.is_ready(c):
                h, p = L.frame_data(SEED, c, pid[c], plen)
                tx.update(c, h, p, MOD_QAM16, FEC_CONV_V27, FEC_NONE)
                pid[c] += 1
        n = min(tx.calls_to_boundary(), ncalls - done)
        out.append(tx.generate(n))
        done += n
    tx.close()
    x = np.concatenate(out) / N
    rx = pkg.MultichannelRx(N, M, cp, taper)
    rx.execute(x)
    fr, pl = rx.poll()
    rx.close()
    same_frames(fr, g["frames"])
    assert np.array_equal(pl, g["payloads"])
