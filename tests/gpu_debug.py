"""first-light diagnostics on the GPU box: oracle vs CUDA on one small multichannel loopback"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refmc import *
import orc
from b2 import pkg

def oracle_channelizer(x, N):
    L = orc.lib()
    K = 2 * N
    q = L.firpfbch_crcf_create_kaiser(0, K, 7, 60.0)
    T = len(x) // K
    dth = np.float32(-0.5) * np.float32(N - 1) / np.float32(N) * np.float32(np.pi)
    u = L.orc_nco_constrain(dth)
    n = np.arange(T * K, dtype=np.uint64)
    th = ((n * np.uint64(u)) & np.uint64(0xffffffff)).astype(np.uint32).astype(np.int32)
    t = (th.astype(np.float64) * (np.pi / 2147483648.0)).astype(np.float32)
    xm = (x[:T * K] * (np.cos(t) - 1j * np.sin(t)).astype(np.complex64)).astype(np.complex64)
    out = np.zeros((T, K), np.complex64)
    y = np.zeros(K, np.complex64)
    for b in range(T):
        xi = np.ascontiguousarray(xm[b * K:(b + 1) * K])
        L.firpfbch_crcf_analyzer_execute(q, xi.ctypes.data, y.ctypes.data)
        out[b] = y
    L.firpfbch_crcf_destroy(q)
    return out[:, :N].T.copy()

def main():
    cfgs = [(8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150),
            (4, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 300),
            (4, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200)]
    Lr = ref_lib()
    for (N, M, cp, tp, mod, f0, f1, plen) in cfgs:
        print("=== config", N, M, cp, tp, mod, f0, f1, plen, flush=True)
        tx = McTx(Lr, N, M, cp, tp)
        ncalls = (M + cp) * 40 * (3 if M == 64 else 1)
        x = tx.run(ncalls, plen, mod, f0, f1, max_frames=2, gain=1.0 / N)
        tx.close()
        rx = McRx(Lr, N, M, cp, tp)
        rx.tap_symbols(True)
        t0 = time.time()
        rx.execute(x)
        t_or = time.time() - t0
        ch_o, idx_o, X_o = rx.symbols()
        fr_o, pl_o = rx.frames()
        rx.close()
        print("oracle frames", len(fr_o), "symbols", len(ch_o), "time %.3f s  %.2f MS/s" % (t_or, len(x) / t_or / 1e6))
        g = pkg.MultichannelRx(N, M, cp, tp)
        g.tap_symbols(True, 1 << 14)
        g.execute(x)
        ms = g.last_timing()
        fr_g, pl_g = g.poll()
        ch_g, idx_g, X_g = g.read_symbols()
        cz = g.read_channelizer()
        print("gpu frames", len(fr_g), "symbols", len(ch_g), "timing ms", ms)
        # channelizer
        co = oracle_channelizer(x, N)
        Tn = min(co.shape[1], cz.shape[1])
        err = np.abs(cz[:, :Tn] - co[:, :Tn]).max() / np.abs(co).max()
        print("channelizer blocks", cz.shape, co.shape, "max rel err %.3e" % err)
        if err > 1e-3:
            bad = np.argwhere(np.abs(cz[:, :Tn] - co[:, :Tn]) > 1e-3 * np.abs(co).max())
            print("first bad entries", bad[:10].tolist())
            d = np.abs(cz[:, :Tn] - co[:, :Tn])
            print("per-channel max err", (d.max(axis=1) / np.abs(co).max()).tolist())
            for (c, b) in bad[:6].tolist():
                print("  ch", c, "blk", b, "gpu", cz[c, b], "orc", co[c, b], "|orc|", abs(co[c, b]), "max", np.abs(co).max())
            print("count bad", len(bad), "of", d.size)
        # frames
        print("oracle:", [(int(f["channel"]), int(f["header_valid"]), int(f["payload_valid"]), int(f["detect_index"]), int(f["complete_index"])) for f in fr_o])
        print("gpu   :", [(int(f["channel"]), int(f["header_valid"]), int(f["payload_valid"]), int(f["detect_index"]), int(f["complete_index"])) for f in fr_g])
        same = len(fr_o) == len(fr_g)
        if same:
            for name in ("channel", "header_valid", "payload_valid", "payload_len", "header", "mod_scheme", "mod_bps", "check", "fec0", "fec1", "detect_index", "complete_index"):
                eq = np.array_equal(fr_o[name], fr_g[name])
                print("  field", name, "equal" if eq else "DIFFERENT")
                same &= eq
            for name in ("evm", "rssi", "cfo"):
                print("  stat", name, "max abs diff %.3e" % np.abs(fr_o[name] - fr_g[name]).max(), fr_o[name][:3], fr_g[name][:3])
            print("  payload bytes equal:", np.array_equal(pl_o, pl_g), len(pl_o), len(pl_g))
        # symbols
        if len(ch_o) and len(ch_g):
            ko = {(int(c), int(i)): k for k, (c, i) in enumerate(zip(ch_o, idx_o))}
            worst = 0.0; matched = 0
            for k, (c, i) in enumerate(zip(ch_g, idx_g)):
                j = ko.get((int(c), int(i)))
                if j is None:
                    continue
                matched += 1
                e = np.abs(X_g[k] - X_o[j]).max() / np.abs(X_o[j]).max()
                worst = max(worst, e)
            print("symbols matched %d / %d (oracle %d), worst rel err %.3e" % (matched, len(ch_g), len(ch_o), worst))
        g.close()
        print("RESULT", "OK" if same else "MISMATCH", flush=True)

if __name__ == "__main__":
    main()
