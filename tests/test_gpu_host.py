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
    assert len(key) > 0 and key == sorted(key)


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


def test_multichanneltx_reset_mid_frame_matches_reference(b2lib):
    """multichanneltx::Reset (lib/multichanneltx.cc:126-149) in the middle of a frame: frame generators and the
    synthesis bank start over, every channel is ready for data again, the NCO keeps its phase (:135) -- the samples
    generated before and after the reset equal the reference class's"""
    N, M, cp, taper = 4, 64, 16, 4
    W = M + cp
    rng = np.random.default_rng(5)
    payloads = [rng.integers(0, 256, 120, dtype=np.uint8) for _ in range(N)]
    out = []
    for lib in (ref_lib(), b2lib):
        tx = McTx(lib, N, M, cp, taper)
        flags = []
        for c in range(N):
            tx.update(c, np.arange(8, dtype=np.uint8) + c, payloads[c], MOD_QPSK, FEC_NONE, FEC_HAMMING128)
        flags.append([tx.is_ready(c) for c in range(N)])
        a = tx.generate(W * 7 + 13)                      # well inside the frame, and not on a symbol boundary
        tx.reset()
        flags.append([tx.is_ready(c) for c in range(N)])
        for c in (1, 3):
            tx.update(c, np.arange(8, dtype=np.uint8) + 10 + c, payloads[c][::-1].copy(), MOD_QAM16, FEC_NONE, FEC_NONE)
        flags.append([tx.is_ready(c) for c in range(N)])
        b = tx.generate(W * 30)
        tx.close()
        out.append((a, b, flags))
    (ao, bo, fo), (ag, bg, fg) = out
    assert fo == fg and fo[1] == [1] * N and fo[0] == [0] * N
    assert np.abs(ag - ao).max() / np.abs(ao).max() < 1e-5
    assert np.abs(bg - bo).max() / np.abs(bo).max() < 1e-5
    # and the frames sent after the reset are whole: the reference receiver decodes them from the CUDA samples
    rx = McRx(ref_lib(), N, M, cp, taper)
    rx.execute(bg)
    fr, pl = rx.frames()
    rx.close()
    assert sorted(fr["channel"].tolist()) == [1, 3] and int(fr["payload_valid"].min()) == 1


DRIVER = os.path.join(ROOT, "tests", "shim", "txrx_driver")


def _read_cf32(path):
    return np.fromfile(path, dtype=np.complex64)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="tests/shim/txrx_driver not built (python __graft_entry__.py)")
def test_ofdmtxrx_split_phase_transmit(tmp_path):
    """assemble_frame / write_symbol / transmit_symbol / end_transmit_frame (lib/ofdmtxrx.cc:366-449) against
    transmit_packet (:297-363): same symbols (here scaled by the caller between write_symbol and transmit_symbol, which
    is what the split API is for), the last buffer sent twice (:347-352 / :431-437), and a receiver decodes both"""
    from b2 import pkg
    a, b = tmp_path / "a.cf32", tmp_path / "b.cf32"
    r = subprocess.run([DRIVER, "split", str(a), str(b)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    nsym = int(r.stdout.split("split-phase symbols:")[1].split()[0])
    M, cp, W = 64, 16, 80
    xa, xb = _read_cf32(a), _read_cf32(b)
    # writesymbol (old API) ends on the last payload symbol, write (new API) appends a tail buffer with the taper's
    # ramp-down: transmit_packet sends nsym + 1 buffers plus the repeated one, the split-phase path nsym plus one
    assert len(xb) == (nsym + 1) * W and len(xa) == (nsym + 2) * W
    body = nsym * W
    # the symbols proper agree (the split path was scaled by 0.5 by the caller); the very first samples of a symbol
    # differ in the taper overlap of the LAST symbol only
    assert np.abs(2.0 * xb[:body - W] - xa[:body - W]).max() / np.abs(xa).max() < 1e-5
    assert np.array_equal(xb[body - W:body], xb[body:body + W])          # the repeated last buffer
    for x, scale in ((xa, 1.0), (xb, 2.0)):
        rx = pkg.OfdmSync(M, cp, 4, streams=1)
        rx.execute(np.concatenate([np.zeros(300, np.complex64), x * scale, np.zeros(600, np.complex64)]))
        fr, pl = rx.poll()
        rx.close()
        assert len(fr) == 1 and int(fr["payload_valid"][0]) == 1 and int(fr["payload_len"][0]) == 200
        assert np.array_equal(pl[:200], (np.arange(200) * 7 + 3).astype(np.uint8))


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="tests/shim/txrx_driver not built (python __graft_entry__.py)")
def test_ofdmtxrx_blocking_receive_worker(tmp_path):
    """ofdmtxrx(..., true) -> ofdmtxrx_rx_worker_blocking (lib/ofdmtxrx.cc:642-739): every received buffer is published
    in *rx_buffer and handed to the synchroniser only after the other thread has edited it.  The capture holds the
    CONJUGATE of a valid transmission; the driver's second thread conjugates every buffer back."""
    from b2 import pkg
    M, cp, taper = 64, 16, 4
    gen = pkg.OfdmGen(M, cp, taper)
    parts = [np.zeros(500, np.complex64)]
    for k in range(8):
        n = gen.assemble(np.arange(8, dtype=np.uint8), (np.arange(200) * 7 + 3).astype(np.uint8), 6, FEC_NONE, FEC_HAMMING128, MOD_QAM16)
        x, _ = gen.write(n + 1)
        parts += [0.5 * x, np.zeros(400, np.complex64)]
    gen.close()
    cap = tmp_path / "conj.cf32"
    np.conj(np.concatenate(parts)).astype(np.complex64).tofile(cap)
    # without the edit nothing decodes ...
    rx = pkg.OfdmSync(M, cp, taper, streams=1)
    rx.execute(_read_cf32(cap))
    fr, _ = rx.poll()
    rx.close()
    assert int(fr["payload_valid"].sum()) == 0 if len(fr) else True
    # ... with it, every packet does
    env = dict(os.environ, B2_UHD_RX_FILE=str(cap))
    r = subprocess.run([DRIVER, "blocking"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    import re
    assert int(re.search(r"valid packets: (\d+)", r.stdout).group(1)) == 8, r.stdout
    assert int(re.search(r"edited buffers: (\d+)", r.stdout).group(1)) >= 1, r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "multichannel_txrx")), reason="reference programs not prebuilt")
def test_reference_multichannel_txrx_program(tmp_path):
    """src/multichannel_txrx.cc UNMODIFIED over the multichanneltxrx class (lib/multichanneltxrx.cc:403-501,541-624) and
    the UHD stand-in: its transmit worker fills a file, its receive worker decodes a looping capture made by
    src/multichannel_tx.cc.  (The program runs for its fixed 30 s.)"""
    air = tmp_path / "air.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(air), B2_UHD_TX_MAX_SAMPLES=str(400000))
    r = subprocess.run([os.path.join(BIN, "multichannel_tx"), "-n", "2"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    out = tmp_path / "txrx_out.cf32"
    env = dict(os.environ, B2_UHD_RX_FILE=str(air), B2_UHD_RX_LOOP="1", B2_UHD_TX_FILE=str(out))
    r = subprocess.run([os.path.join(BIN, "multichannel_txrx"), "-q"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    import re
    m = re.search(r"valid packets\s*:\s*(\d+)", r.stdout)
    assert m and int(m.group(1)) >= 10, r.stdout[-1500:]
    assert "transmitting packet" in r.stdout and out.stat().st_size > 100000
