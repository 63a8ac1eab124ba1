"""End-to-end pins for the CPU oracle: the reference's own lib/multichanneltx.cc ->
lib/multichannelrx.cc (compiled unmodified into oracle/_ref/libref_mc.so over the oracle's
liquid API) must decode every frame it sends, independent of how the stream is chunked
(SURVEY.md Q4), on the BASELINE.json config shapes (scaled down in frame count)."""
import numpy as np
import pytest

from refmc import (McRx, McTx, ref_lib, payload_of, FEC_NONE, FEC_HAMMING128, FEC_CONV_V27,
                   MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256)

SEED = 0xB2000000

CASES = [
    # N, M, cp, taper, mod, fec0, fec1, payload_len
    (1, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_NONE, 200),            # C1 shape
    (8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150),      # C2 shape
    (4, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 300),      # C3 shape (fewer channels)
    (2, 512, 64, 16, MOD_QAM256, FEC_NONE, FEC_NONE, 1200),       # C4 modulation / M
    (4, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200),        # C5 shape (fewer channels)
    (3, 48, 12, 3, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 64),       # non power-of-two N and M
]


def frame_symbols(M, mod_bps, enc_len):
    from orc import default_sctype
    nd = int((default_sctype(M) == 2).sum())
    nh = -(-288 // nd)
    npay = -(-(-(-8 * enc_len // mod_bps)) // nd)
    return 3 + nh + npay + 1          # + tail buffer (write() interface)


def run_loopback(N, M, cp, taper, mod, fec0, fec1, plen, nframes=3, chunk=0):
    L = ref_lib()
    import orc
    enc = orc.lib().orc_packetizer_enc_len(plen, 6, fec0, fec1)
    bps = {MOD_QPSK: 2, MOD_QAM16: 4, MOD_QAM64: 6, MOD_QAM256: 8}[mod]
    nsym = frame_symbols(M, bps, enc)
    ncalls = (nsym * nframes + 4) * (M + cp)
    tx = McTx(L, N, M, cp, taper)
    x = tx.run(ncalls, plen, mod, fec0, fec1, seed=SEED, max_frames=nframes, gain=1.0 / N)
    tx.close()
    rx = McRx(L, N, M, cp, taper)
    rx.execute(x, chunk)
    fr, pl = rx.frames()
    rx.close()
    return x, fr, pl, nsym


@pytest.mark.parametrize("case", CASES)
def test_loopback_decodes_every_frame(case):
    N, M, cp, taper, mod, fec0, fec1, plen = case
    nframes = 3
    x, fr, pl, nsym = run_loopback(*case, nframes=nframes)
    assert len(fr) == N * nframes
    L = ref_lib()
    for i in range(len(fr)):
        c = int(fr["channel"][i])
        pid = int(fr["header"][i][0]) * 256 + int(fr["header"][i][1])
        h, p = L.frame_data(SEED, c, pid, plen)
        assert fr["header_valid"][i] == 1 and fr["payload_valid"][i] == 1
        assert np.array_equal(fr["header"][i], h)
        assert np.array_equal(payload_of(fr, pl, i), p)
        assert (fr["mod_scheme"][i], fr["fec0"][i], fr["fec1"][i], fr["check"][i]) == (mod, fec0, fec1, 6)
        assert fr["evm"][i] < -30.0
    # callback order: ascending completion time, then channel
    key = list(zip(fr["complete_index"].tolist(), fr["channel"].tolist()))
    assert key == sorted(key)
    # frames of one channel complete one frame period apart (+-1 from the timing estimate) (the detect index rides on
    # the SEEK grid, which restarts after every frame, so it is only within M of periodic)
    for c in range(N):
        sel = fr["channel"] == c
        assert np.all(np.abs(np.diff(fr["complete_index"][sel].astype(np.int64)) - nsym * (M + cp)) <= 2)
        assert np.all(np.abs(np.diff(fr["detect_index"][sel].astype(np.int64)) - nsym * (M + cp)) < M)


def test_chunking_invariance():
    case = CASES[1]
    x, fr0, pl0, _ = run_loopback(*case, nframes=2)
    for chunk in (1, 7, 16, 1000):
        _, fr, pl, _ = run_loopback(*case, nframes=2, chunk=chunk)
        assert fr.tobytes() == fr0.tobytes()
        assert pl.tobytes() == pl0.tobytes()


def test_idle_channels_and_noise_floor_do_not_false_alarm():
    N, M, cp, taper = 4, 64, 16, 4
    L = ref_lib()
    tx = McTx(L, N, M, cp, taper)
    x = tx.run(80 * 40, 50, MOD_QPSK, FEC_NONE, FEC_NONE, seed=SEED, channel_mask=0b0101, max_frames=1, gain=0.25)
    tx.close()
    rng = np.random.default_rng(11)
    x = x + (1e-3 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
    rx = McRx(L, N, M, cp, taper)
    rx.execute(x)
    fr, pl = rx.frames()
    assert sorted(fr["channel"].tolist()) == [0, 2]
    assert np.all(fr["payload_valid"] == 1)


def test_corrupted_payload_reports_invalid():
    N, M, cp, taper = 1, 64, 16, 4
    L = ref_lib()
    tx = McTx(L, N, M, cp, taper)
    x = tx.run(80 * 30, 100, MOD_QPSK, FEC_NONE, FEC_NONE, seed=SEED, max_frames=1)
    tx.close()
    # wipe one payload symbol (wideband samples: K=2 per channel sample)
    s = (3 + 7 + 4) * 80 * 2
    x[s:s + 160] = 0
    rx = McRx(L, N, M, cp, taper)
    rx.execute(x)
    fr, _ = rx.frames()
    assert len(fr) == 1 and fr["header_valid"][0] == 1 and fr["payload_valid"][0] == 0


def test_constructor_errors_throw():
    L = ref_lib()
    for args in ((0, 64, 16, 4), (2, 6, 2, 0), (2, 64, 0, 0), (2, 64, 4, 8)):
        with pytest.raises(ValueError):
            McRx(L, *args)
        with pytest.raises(ValueError):
            McTx(L, *args)
