"""Known-answer and property tests that pin the CPU oracle (oracle/).  The reference ships no
tests or golden vectors for this path, so the pins are external
known answers (CRC-32 check value, code distances, scipy/numpy cross-checks) and algebraic
properties."""
import itertools
import ctypes as C

import numpy as np
import pytest
import scipy.signal
import scipy.special

import orc
from refmc import FEC_NONE, FEC_HAMMING128, FEC_GOLAY2412, FEC_CONV_V27, CRC_32, MOD_BPSK, MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256


def test_crc32_check_value():
    assert orc.crc32(np.frombuffer(b"123456789", np.uint8)) == 0xCBF43926
    import zlib
    rng = np.random.default_rng(1)
    for n in (1, 2, 14, 100, 1200):
        m = rng.integers(0, 256, n, dtype=np.uint8)
        assert orc.crc32(m) == zlib.crc32(m.tobytes())


def test_hamming128_corrects_every_single_error():
    L = orc.lib()
    codes = [L.orc_hamming128_encode_symbol(s) for s in range(256)]
    assert len(set(codes)) == 256
    dmin = min(bin(a ^ b).count("1") for a, b in itertools.combinations(codes, 2))
    assert dmin == 3
    for s in range(256):
        c = codes[s]
        assert L.orc_hamming128_decode_symbol(c) == s
        for b in range(12):
            assert L.orc_hamming128_decode_symbol(c ^ (1 << b)) == s


def test_golay2412_corrects_up_to_three_errors():
    L = orc.lib()
    rng = np.random.default_rng(2)
    msgs = list(rng.integers(0, 4096, 40)) + [0, 1, 0xfff, 0x800]
    for s in msgs:
        s = int(s)
        c = L.orc_golay2412_encode_symbol(s)
        assert c & 0xfff == s
        assert L.orc_golay2412_decode_symbol(c) == s
        for nerr in (1, 2, 3):
            for pos in itertools.islice(itertools.combinations(range(24), nerr), 0, None, 7 if nerr == 3 else 1):
                e = 0
                for p in pos:
                    e |= 1 << p
                assert L.orc_golay2412_decode_symbol(c ^ e) == s
    w = [bin(L.orc_golay2412_encode_symbol(s)).count("1") for s in range(1, 4096)]
    assert min(w) == 8


@pytest.mark.parametrize("scheme", [FEC_NONE, FEC_HAMMING128, FEC_GOLAY2412, FEC_CONV_V27])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 18, 100, 1204])
def test_fec_roundtrip(scheme, n):
    rng = np.random.default_rng(n)
    m = rng.integers(0, 256, n, dtype=np.uint8)
    e = orc.fec_encode(scheme, m)
    assert np.array_equal(orc.fec_decode(scheme, n, e), m)


def test_conv27_corrects_scattered_errors():
    rng = np.random.default_rng(3)
    m = rng.integers(0, 256, 200, dtype=np.uint8)
    e = orc.fec_encode(FEC_CONV_V27, m)
    assert len(e) == (2 * (8 * 200 + 6) + 7) // 8
    e2 = e.copy()
    for byte in range(5, len(e2), 13):      # one bit error every 104 coded bits
        e2[byte] ^= 1 << (byte % 8)
    assert np.array_equal(orc.fec_decode(FEC_CONV_V27, 200, e2), m)
    z = orc.fec_encode(FEC_CONV_V27, np.zeros(4, np.uint8))
    assert not z.any()
    # impulse response = the two generator polynomials 0x6d / 0x4f, MSB first
    one = orc.fec_encode(FEC_CONV_V27, np.array([0x80], np.uint8))
    bits = np.unpackbits(one)[:14].reshape(7, 2)
    ga = [(0x6d >> i) & 1 for i in range(7)]
    gb = [(0x4f >> i) & 1 for i in range(7)]
    assert list(bits[:, 0]) == ga and list(bits[:, 1]) == gb


def test_conv27_against_an_independent_numpy_decoder():
    """the oracle's r1/2 K=7 codec against the textbook numpy model in tools/vit_merge_stats.py (state = shift register,
    Hamming branch metrics, strict-less tie-break): same code bits, same decoded bytes under 2 % channel bit errors"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import vit_merge_stats as V
    rng = np.random.default_rng(17)
    for n in (5, 64, 150):
        m = rng.integers(0, 256, n, dtype=np.uint8)
        nbits = 8 * n + 6
        msgbits = np.concatenate([np.unpackbits(m), np.zeros(6, np.uint8)]).astype(np.int64)
        pairs = V.encode(msgbits)                                 # (c(0x6d) << 1) | c(0x4f) per input bit
        e = orc.fec_encode(FEC_CONV_V27, m)
        ebits = np.unpackbits(e)[:2 * nbits].reshape(nbits, 2)
        assert np.array_equal((ebits[:, 0].astype(np.int64) << 1) | ebits[:, 1], pairs)
        flips = (rng.random(2 * nbits) < 0.02).astype(np.uint8)
        rxbits = np.unpackbits(e).copy()
        rxbits[:2 * nbits] ^= flips
        dec_oracle = orc.fec_decode(FEC_CONV_V27, n, np.packbits(rxbits))
        rx = (rxbits[:2 * nbits].reshape(nbits, 2)[:, 0].astype(np.int64) << 1) | rxbits[:2 * nbits].reshape(nbits, 2)[:, 1]
        start = np.full(V.NS, 63, np.int64); start[0] = 0
        dec, _ = V.acs(rx, 0, nbits, start)
        bits = V.traceback(dec, 0, nbits, 0, 0)[:8 * n]
        assert np.array_equal(np.packbits(bits), dec_oracle)
        assert np.array_equal(dec_oracle, m)


@pytest.mark.parametrize("n", [2, 3, 16, 36, 100, 255, 1204, 1806, 2410])
def test_interleaver_is_a_bit_permutation_and_inverts(n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 256, n, dtype=np.uint8)
    y = orc.interleave(x, 4, False)
    assert np.array_equal(orc.interleave(y, 4, True), x)
    assert np.unpackbits(x).sum() == np.unpackbits(y).sum()
    if n <= 100:
        seen = set()
        for b in range(8 * n):
            e = np.zeros(n, np.uint8)
            e[b // 8] = 0x80 >> (b % 8)
            o = np.flatnonzero(np.unpackbits(orc.interleave(e, 4, False)))
            assert len(o) == 1
            seen.add(int(o[0]))
        assert len(seen) == 8 * n
    assert np.array_equal(orc.interleave(x, 0, False), x)


@pytest.mark.parametrize("fec0,fec1", [(FEC_NONE, FEC_NONE), (FEC_NONE, FEC_HAMMING128), (FEC_CONV_V27, FEC_NONE),
                                       (FEC_GOLAY2412, FEC_NONE), (FEC_CONV_V27, FEC_HAMMING128)])
def test_packetizer_roundtrip_and_crc_detects(fec0, fec1):
    rng = np.random.default_rng(5)
    for n in (1, 14, 100, 1200):
        m = rng.integers(0, 256, n, dtype=np.uint8)
        pkt = orc.packetizer_encode(m, CRC_32, fec0, fec1)
        assert len(pkt) == orc.lib().orc_packetizer_enc_len(n, CRC_32, fec0, fec1)
        d, ok = orc.packetizer_decode(pkt, n, CRC_32, fec0, fec1)
        assert ok and np.array_equal(d, m)
        if fec0 == FEC_NONE and fec1 == FEC_NONE:
            pkt[0] ^= 0x10
            d, ok = orc.packetizer_decode(pkt, n, CRC_32, fec0, fec1)
            assert not ok
    L = orc.lib()
    assert L.orc_packetizer_enc_len(1200, CRC_32, FEC_NONE, FEC_HAMMING128) == 1806
    assert L.orc_packetizer_enc_len(1200, CRC_32, FEC_CONV_V27, FEC_NONE) == 2410
    assert L.orc_packetizer_enc_len(14, CRC_32, FEC_GOLAY2412, FEC_NONE) == 36


def test_kaiser_design_matches_scipy():
    L = orc.lib()
    assert abs(L.orc_kaiser_beta_As(60.0) - scipy.signal.kaiser_beta(60.0)) < 1e-5
    for K, m in ((16, 7), (128, 13), (512, 7)):
        n = 2 * K * m + 1
        h = orc.firdes_kaiser(n, 0.5 / K, 60.0)
        t = np.arange(n) - (n - 1) / 2
        beta = 0.1102 * (60.0 - 8.7)
        r = 2 * t / n                                   # liquid's kaiser() divides by n, not n-1
        w = scipy.special.i0(beta * np.sqrt(1 - r * r)) / scipy.special.i0(beta)
        ref = np.sinc(2 * (0.5 / K) * t) * w
        assert np.max(np.abs(h - ref)) < 2e-5
        assert abs(h[(n - 1) // 2] - 1.0) < 1e-6 and np.allclose(h, h[::-1], atol=1e-6)


@pytest.mark.parametrize("n", [2, 8, 16, 64, 128, 512, 1024, 6, 48])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    X = orc.fft(x)
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.max(np.abs(X - ref)) / np.max(np.abs(ref)) < 2e-6
    xb = orc.fft(x, backward=True)
    refb = np.fft.ifft(x.astype(np.complex128)) * n
    assert np.max(np.abs(xb - refb)) / np.max(np.abs(refb)) < 2e-6


@pytest.mark.parametrize("scheme,bps", [(MOD_BPSK, 1), (MOD_QPSK, 2), (MOD_QAM16, 4), (MOD_QAM64, 6), (MOD_QAM256, 8)])
def test_modem_unit_energy_and_gray(scheme, bps):
    L = orc.lib()
    L.orc_modem_modulate.restype = C.c_double   # float _Complex comes back as two packed floats in xmm0
    q = orc.Modem()
    L.orc_modem_init(C.byref(q), scheme)
    assert q.bps == bps
    pts = []
    for s in range(1 << bps):
        d = L.orc_modem_modulate(C.byref(q), s)
        re, im = np.frombuffer(np.float64(d).tobytes(), np.float32)
        pts.append(complex(re, im))
    pts = np.array(pts)
    assert abs(np.mean(np.abs(pts) ** 2) - 1.0) < 1e-5
    assert len(set(np.round(pts, 5))) == 1 << bps
    dmin = np.min(np.abs(pts[:, None] - pts[None, :]) + 10 * np.eye(len(pts)))
    for a in range(len(pts)):
        for b in range(len(pts)):
            if a != b and abs(abs(pts[a] - pts[b]) - dmin) < 1e-5:
                assert bin(a ^ b).count("1") == 1


def test_msequence_periods():
    L = orc.lib()
    for m in range(2, 12):
        ms = orc.Mseq()
        L.orc_mseq_init_default(C.byref(ms), m)
        n = (1 << m) - 1
        bits = [L.orc_mseq_advance(C.byref(ms)) for _ in range(2 * n)]
        assert bits[:n] == bits[n:]
        assert sum(bits[:n]) == (1 << (m - 1))          # balance property of an m-sequence
        for p in range(1, n):
            if n % p == 0:
                assert bits[:n - p] != bits[p:n]


def test_default_allocation_counts():
    for M, (nd, npil, nn) in {64: (44, 6, 14), 256: (178, 26, 52), 512: (356, 52, 104), 48: (34, 4, 10)}.items():
        p = orc.default_sctype(M)
        assert ((p == 2).sum(), (p == 1).sum(), (p == 0).sum()) == (nd, npil, nn)
        assert p[0] == 0


def test_nco_offset_is_exact_fixed_point():
    L = orc.lib()
    assert L.orc_nco_constrain(0.0) == 0
    for N in (2, 8, 64, 256):
        off = np.float32(-0.5) * np.float32(N - 1) / np.float32(N) * np.float32(np.pi)
        u = L.orc_nco_constrain(off)
        want = (-(N - 1) / (4 * N)) % 1.0
        assert abs(u / 2 ** 32 - want) < 1e-7


def test_channelizer_synthesis_then_analysis_recovers_channels():
    L = orc.lib()
    K = 16
    syn = L.firpfbch_crcf_create_kaiser(1, K, 13, 60.0)
    ana = L.firpfbch_crcf_create_kaiser(0, K, 7, 60.0)
    rng = np.random.default_rng(7)
    T = 400
    X = np.zeros((T, K), np.complex64)
    # drive the lower half of the channels only, as multichanneltx does (lib/multichanneltx.cc:205-210)
    X[:, :K // 2] = (rng.standard_normal((T, K // 2)) + 1j * rng.standard_normal((T, K // 2))) * 0.5
    lp = scipy.signal.firwin(63, 0.08)   # narrow-band: inside the flat part of the channel response
    X[:, :K // 2] = scipy.signal.lfilter(lp, 1.0, X[:, :K // 2], axis=0).astype(np.complex64)
    y = np.zeros(K, np.complex64)
    Y = np.zeros((T, K), np.complex64)
    for t in range(T):
        xin = np.ascontiguousarray(X[t])
        L.firpfbch_crcf_synthesizer_execute(syn, xin.ctypes.data, y.ctypes.data)
        yy = np.zeros(K, np.complex64)
        L.firpfbch_crcf_analyzer_execute(ana, y.ctypes.data, yy.ctypes.data)
        Y[t] = yy
    L.firpfbch_crcf_destroy(syn)
    L.firpfbch_crcf_destroy(ana)
    for c in (0, 3, K // 2 - 1):
        a, b = X[:, c], Y[:, c]
        corr = [np.vdot(a[:T - d], b[d:]) for d in range(40)]
        d = int(np.argmax(np.abs(corr)))
        g = corr[d] / np.vdot(a[:T - d], a[:T - d])
        err = b[d:] - g * a[:T - d]
        snr = 10 * np.log10(np.sum(np.abs(g * a[50:T - d]) ** 2) / np.sum(np.abs(err[50:]) ** 2))
        assert snr > 35, (c, d, snr)


def test_msresamp_tone_and_chunking():
    n = 4000
    f = 0.05
    x = np.exp(2j * np.pi * f * np.arange(n)).astype(np.complex64)
    rate = np.float32(1.07)
    y = orc.msresamp(x, rate)
    assert abs(len(y) - n * 1.07) <= 2
    k = np.arange(200, len(y) - 200)
    step = round(2 ** 32 / float(rate)) / 2 ** 32
    ref = np.exp(2j * np.pi * f * (k * step - 7))
    assert np.max(np.abs(y[k] - ref)) < 2e-3
    L = orc.lib()
    q = L.msresamp_crcf_create(rate, 60.0)
    out = []
    i = 0
    rng = np.random.default_rng(0)
    while i < n:
        c = int(rng.integers(1, 97))
        xi = np.ascontiguousarray(x[i:i + c])
        yo = np.zeros(2 * len(xi) + 8, np.complex64)
        ny = C.c_uint(0)
        L.msresamp_crcf_execute(q, xi.ctypes.data, len(xi), yo.ctypes.data, C.byref(ny))
        out.append(yo[:ny.value])
        i += c
    L.msresamp_crcf_destroy(q)
    assert np.array_equal(np.concatenate(out), y)
