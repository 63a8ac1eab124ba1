#!/usr/bin/env python
"""Check the CUDA path against the committed golden fixtures with NOTHING of the oracle in the process: only numpy and
the product library are loaded (asserted from /proc/self/maps at the end).  Run by tests/test_golden_gpu.py in a
subprocess on the GPU box; prints "golden ok" on success.

  c3_n2_m256_v27       multichannelrx, 2 channels, M=256, 16-QAM, conv r1/2 K=7      (BASELINE configs[2], scaled)
  c5_n2_m512_qam64     multichannelrx, 2 channels, M=512, 64-QAM, 1200-byte payloads (BASELINE configs[4], scaled)
  c4_link_m512_qam256  msresamp 1/1.07 -> ofdmflexframesync, M=512, 256-QAM          (BASELINE configs[3], scaled)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("liquid-usrp_b200")

EXACT = ("channel", "header_valid", "payload_valid", "payload_len", "header", "mod_scheme", "mod_bps",
         "check", "fec0", "fec1", "detect_index", "complete_index", "payload_offset")


def same(fr, pl, g):
    assert len(fr) == len(g["frames"]), (len(fr), len(g["frames"]))
    for k in EXACT:
        assert np.array_equal(fr[k], g["frames"][k]), k
    assert np.array_equal(pl, g["payloads"])
    np.testing.assert_allclose(fr["evm"], g["frames"]["evm"], atol=2e-3)
    np.testing.assert_allclose(fr["cfo"], g["frames"]["cfo"], atol=1e-6)


def main():
    g = np.load(os.path.join(HERE, "c3_n2_m256_v27.npz"))
    rx = pkg.MultichannelRx(2, 256, 32, 8)
    rx.execute(g["x"])
    fr, pl = rx.poll()
    rx.close()
    same(fr, pl, g)
    g = np.load(os.path.join(HERE, "c5_n2_m512_qam64.npz"))
    rx = pkg.MultichannelRx(2, 512, 64, 16)
    rx.tap_symbols(True, 4096)
    cut = len(g["x"]) // 3 + 5                       # ragged calls
    rx.execute(g["x"][:cut])
    rx.execute(g["x"][cut:])
    fr, pl = rx.poll()
    ch, idx, X = rx.read_symbols()
    rx.close()
    same(fr, pl, g)
    sel = ch == 0
    assert np.array_equal(idx[sel], g["sym_index"])
    err = np.abs(X[sel] - g["sym_X"]).max(axis=1) / np.abs(g["sym_X"]).max(axis=1)
    assert err.max() < 1e-5, err.max()               # equalised symbols: 1e-5 relative (north star)
    g = np.load(os.path.join(HERE, "c4_link_m512_qam256.npz"))
    down = pkg.MsResamp(np.float32(1.0) / g["rate"])
    z = down.execute(g["x"])
    down.close()
    rx = pkg.OfdmSync(512, 64, 16, streams=1, max_batch=len(z))
    rx.execute(z.reshape(1, -1))
    fr, pl = rx.poll()
    rx.close()
    same(fr, pl, g)
    maps = open("/proc/self/maps").read()
    assert "liborc" not in maps and "libref_mc" not in maps and "libb200ofdm.so" in maps
    print("golden ok")


if __name__ == "__main__":
    main()
