"""Parity of the CUDA receive path (through the C ABI, include/b200_ofdm.h) against the CPU
oracle -- the reference's lib/multichanneltx.cc / lib/multichannelrx.cc compiled over oracle/.

Bars (BASELINE.json north_star): decoded payload bytes and frame detect/complete sample indices
bit-exact; equalised symbols within 1e-5 relative (max |X_gpu - X_ref| / max |X_ref| per OFDM
symbol); frame stats (evm/rssi in dB, cfo) to 1e-3."""
import os
import numpy as np
import pytest

from refmc import (McRx, McTx, ref_lib, payload_of, FEC_NONE, FEC_HAMMING128, FEC_GOLAY2412, FEC_CONV_V27,
                   MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256)

pytestmark = pytest.mark.gpu
SEED = 0xB2000000
EXACT = ("channel", "header_valid", "payload_valid", "payload_len", "header", "mod_scheme", "mod_bps",
         "check", "fec0", "fec1", "detect_index", "complete_index")

CASES = {
    # name: N, M, cp, taper, mod, fec0, fec1, payload_len, frames, noise_std
    "c1_loopback_1ch": (1, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_NONE, 200, 3, 0.0),
    "c2_8ch_h128": (8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150, 3, 0.0),
    "c3_16ch_qam16_v27": (16, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 300, 2, 0.0),
    "c4_shape_qam256": (2, 512, 64, 16, MOD_QAM256, FEC_NONE, FEC_NONE, 1200, 2, 0.0),
    "c5_shape_32ch_qam64": (32, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200, 2, 0.0),
    "golay_outer_h128": (4, 128, 16, 4, MOD_QAM16, FEC_GOLAY2412, FEC_HAMMING128, 97, 2, 0.0),
    "noisy_30dB": (8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150, 3, 0.03),
    "noisy_v27": (4, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 200, 2, 0.05),
    # the register-resident synchroniser's other FFT plans: 4 passes (8,8,8,2 / 8,8,8,4 / 8^4)
    "m1024_2ch_qam16_h128": (2, 1024, 128, 32, MOD_QAM16, FEC_NONE, FEC_HAMMING128, 900, 2, 0.0),
    "m2048_1ch_qpsk": (1, 2048, 64, 16, MOD_QPSK, FEC_NONE, FEC_NONE, 700, 2, 0.0),
    "m4096_1ch_qam64": (1, 4096, 256, 64, MOD_QAM64, FEC_NONE, FEC_NONE, 3000, 2, 0.0),
    "m256_qam256_golay": (4, 256, 32, 8, MOD_QAM256, FEC_GOLAY2412, FEC_NONE, 500, 3, 0.0),
    # sizes that are not powers of two: the reference programs default to M = 48, cp 6, taper 4, one channel
    # (src/multichannel_rx.cc:88-95); channel counts 3, 5, 6, 7 give K = 6, 10, 12, 14 point filterbanks
    "ref_defaults_m48_1ch": (1, 48, 6, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 100, 3, 0.0),
    "n3_m48_qam16": (3, 48, 6, 4, MOD_QAM16, FEC_NONE, FEC_NONE, 120, 2, 0.0),
    "n6_m80_h128": (6, 80, 10, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 100, 2, 0.0),
    "n5_m96_v27": (5, 96, 12, 4, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 150, 2, 0.0),
    "n7_m120_qam64": (7, 120, 12, 4, MOD_QAM64, FEC_NONE, FEC_NONE, 200, 2, 0.0),
    "n17_m58_qpsk": (17, 58, 8, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 100, 2, 0.0),
    # BASELINE configs[4] at its full width (the oracle receiver needs ~1 s for two frames per channel)
    "c5_full_256ch_qam64": (256, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200, 2, 0.0),
    "c5_full_256ch_qam64_30dB": (256, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200, 2, 0.03),
    # BASELINE configs[2] at its full width: 64 channels, M = 256, 16-QAM, conv r1/2 K=7
    "c3_full_64ch_qam16_v27": (64, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 1200, 2, 0.0),
}


def make_input(case, seed=1):
    N, M, cp, taper, mod, fec0, fec1, plen, nframes, noise = case
    import orc
    from test_oracle_loopback import frame_symbols
    bps = {MOD_QPSK: 2, MOD_QAM16: 4, MOD_QAM64: 6, MOD_QAM256: 8}[mod]
    nsym = frame_symbols(M, bps, orc.lib().orc_packetizer_enc_len(plen, 6, fec0, fec1))
    ncalls = (nsym * nframes + 3) * (M + cp) + 7
    tx = McTx(ref_lib(), N, M, cp, taper)
    x = tx.run(ncalls, plen, mod, fec0, fec1, seed=SEED, max_frames=nframes, gain=1.0 / N)
    tx.close()
    if noise:
        rng = np.random.default_rng(seed)
        s = noise * np.sqrt(np.mean(np.abs(x) ** 2))
        x = (x + s * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) / np.sqrt(2)).astype(np.complex64)
    return x


def run_oracle(case, x, tap=False):
    N, M, cp, taper = case[:4]
    rx = McRx(ref_lib(), N, M, cp, taper)
    if tap:
        rx.tap_symbols(True)
    rx.execute(x)
    sym = rx.symbols() if tap else None
    fr, pl = rx.frames()
    rx.close()
    return fr, pl, sym


def run_gpu(case, x, chunks=None, tap=False, max_batch=0):
    from b2 import pkg
    N, M, cp, taper = case[:4]
    g = pkg.MultichannelRx(N, M, cp, taper, max_batch=max_batch)
    if tap:
        g.tap_symbols(True, 1 << 15)
    frs, pls = [], []
    if chunks is None:
        chunks = [len(x)]
    i = 0
    for c in chunks:
        g.execute(x[i:i + c])
        i += c
    assert i == len(x)
    fr, pl = g.poll()
    sym = g.read_symbols() if tap else None
    g.close()
    return fr, pl, sym


def assert_frames_equal(fo, po, fg, pg):
    assert len(fg) == len(fo), (len(fg), len(fo))
    for name in EXACT:
        assert np.array_equal(fo[name], fg[name]), name
    assert np.array_equal(po, pg)
    assert np.array_equal(fo["payload_offset"], fg["payload_offset"])
    np.testing.assert_allclose(fg["evm"], fo["evm"], atol=2e-3)
    np.testing.assert_allclose(fg["rssi"], fo["rssi"], atol=1e-3)
    np.testing.assert_allclose(fg["cfo"], fo["cfo"], atol=1e-6)


@pytest.mark.parametrize("name", list(CASES))
def test_frames_bit_exact_and_symbols_within_tolerance(name):
    case = CASES[name]
    x = make_input(case)
    fo, po, so = run_oracle(case, x, tap=True)
    fg, pg, sg = run_gpu(case, x, tap=True)
    assert len(fo) == case[0] * case[8]
    assert int(fo["payload_valid"].sum()) == len(fo)
    assert_frames_equal(fo, po, fg, pg)
    # equalised symbols, matched by (channel, sample index)
    cho, io, Xo = so
    chg, ig, Xg = sg
    assert len(cho) == len(chg) and len(cho) > 0
    oo = np.lexsort((io, cho))
    og = np.lexsort((ig, chg))
    assert np.array_equal(cho[oo], chg[og]) and np.array_equal(io[oo], ig[og])
    err = np.abs(Xg[og] - Xo[oo]).max(axis=1) / np.abs(Xo[oo]).max(axis=1)
    tol = 1e-5 if case[9] == 0.0 else 2e-5
    assert err.max() < tol, err.max()


def test_chunking_invariance_on_device():
    case = CASES["c2_8ch_h128"]
    x = make_input(case)
    fo, po, _ = run_oracle(case, x)
    rng = np.random.default_rng(5)
    for pattern in ("odd", "tiny", "blocks", "small_batch"):
        if pattern == "odd":
            chunks, left = [], len(x)
            while left:
                c = int(min(left, rng.integers(1, 5000)))
                chunks.append(c)
                left -= c
            fg, pg, _ = run_gpu(case, x, chunks)
        elif pattern == "tiny":
            chunks = [1] * 40 + [3] * 11 + [len(x) - 73]
            fg, pg, _ = run_gpu(case, x, chunks)
        elif pattern == "blocks":
            chunks = [16 * 100] * (len(x) // 1600) + ([len(x) % 1600] if len(x) % 1600 else [])
            fg, pg, _ = run_gpu(case, x, chunks)
        else:
            fg, pg, _ = run_gpu(case, x, None, max_batch=4096)       # internal splitting of one call
        assert_frames_equal(fo, po, fg, pg)


def test_chunking_invariance_of_the_production_kernels():
    """the register-resident synchroniser (pipelined payload symbols, chunk-boundary state) and the
    column-per-thread channelizer under ragged calls: 1 .. 30000 wideband samples at a time"""
    for name in ("c5_shape_32ch_qam64", "c3_full_64ch_qam16_v27"):
        case = CASES[name]
        x = make_input(case)
        fo, po, _ = run_oracle(case, x)
        rng = np.random.default_rng(11)
        chunks, left = [], len(x)
        while left:
            c = int(min(left, rng.integers(1, 30000)))
            chunks.append(c)
            left -= c
        fg, pg, _ = run_gpu(case, x, chunks)
        assert_frames_equal(fo, po, fg, pg)
        # batches far shorter than a frame: every frame completes in a batch it did not begin in
        fg, pg, _ = run_gpu(case, x, None, max_batch=5000)
        assert_frames_equal(fo, po, fg, pg)


@pytest.mark.parametrize("name,frac", [("c2_8ch_h128", 0.34), ("c5_shape_32ch_qam64", 0.3), ("c5_shape_32ch_qam64", 0.62),
                                       ("c3_16ch_qam16_v27", 0.45)])
def test_reset_mid_stream_matches_oracle(name, frac):
    """Reset() in the middle of a stream, also in the middle of a frame and with the register-resident
    synchroniser's worker pairs in whatever roles they happen to hold (the M >= 256 cases)"""
    case = CASES[name]
    N, M, cp, taper = case[:4]
    x = make_input(case)
    cut = int(len(x) * frac) + 5
    from b2 import pkg
    rx = McRx(ref_lib(), N, M, cp, taper)
    rx.execute(x[:cut]); rx.reset(); rx.execute(x[cut:])
    fo, po = rx.frames()
    rx.close()
    g = pkg.MultichannelRx(N, M, cp, taper)
    g.execute(x[:cut]); g.reset(); g.execute(x[cut:])
    fg, pg = g.poll()
    g.close()
    assert_frames_equal(fo, po, fg, pg)


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_conv_decode_modes_with_corrupted_frames(mode, monkeypatch):
    """conv-coded frames through the exact decoder (0), the speculative decode (1: segmented recursion, a launch with few
    frames) and the thread-parallel speculative traceback alone (2: what a launch with more frames than CTAs takes);
    clean, noisy and CRC-failing frames -- the failing ones fall back to the exact decoder, so every payload byte,
    valid or not, equals the oracle's"""
    monkeypatch.setenv("B2_VIT_MODE", mode)
    nbad = 0
    for name in ("c3_16ch_qam16_v27", "noisy_v27", "n5_m96_v27"):
        case = CASES[name]
        x = make_input(case).copy()
        N = case[0]
        # wipe part of the payloads: a long gap (CRC fails on several channels) and a short one (the Viterbi decoder
        # corrects most of it)
        p0 = int(len(x) * 0.35)
        x[p0:p0 + N * 200] = 0
        p0 = int(len(x) * 0.8)
        x[p0:p0 + N * 100] = 0
        fo, po, _ = run_oracle(case, x)
        fg, pg, _ = run_gpu(case, x)
        assert len(fo) > 0 and int(fo["header_valid"].sum()) > 0
        nbad += int(((fo["header_valid"] == 1) & (fo["payload_valid"] == 0)).sum())
        assert_frames_equal(fo, po, fg, pg)
    assert nbad >= 3            # the fallback was exercised


def test_idle_channels_corruption_and_noise_only():
    N, M, cp, taper = 4, 64, 16, 4
    case = (N, M, cp, taper)
    tx = McTx(ref_lib(), N, M, cp, taper)
    x = tx.run(80 * 60, 50, MOD_QPSK, FEC_NONE, FEC_NONE, seed=SEED, channel_mask=0b0101, max_frames=2, gain=0.25)
    tx.close()
    rng = np.random.default_rng(11)
    x = x + (1e-3 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
    # wipe part of one payload so the CRC fails on some frames
    x[17000:17400] = 0
    fo, po, _ = run_oracle(case, x)
    fg, pg, _ = run_gpu(case, x)
    assert sorted(set(fo["channel"].tolist())) == [0, 2]
    assert int((fo["payload_valid"] == 0).sum()) >= 1
    assert_frames_equal(fo, po, fg, pg)
    # all-zero and noise-only input: no frames, no NaN trouble
    z = np.zeros(80 * 40 * 8, np.complex64)
    fg, pg, _ = run_gpu(case, z)
    assert len(fg) == 0
    n = (0.1 * (rng.standard_normal(len(z)) + 1j * rng.standard_normal(len(z)))).astype(np.complex64)
    fo, po, _ = run_oracle(case, n)
    fg, pg, _ = run_gpu(case, n)
    assert_frames_equal(fo, po, fg, pg)


def test_channelizer_output_matches_oracle():
    import orc
    from b2 import pkg
    for N in (1, 3, 5, 7, 8, 11, 12, 17, 31, 32, 64, 128, 256):      # K = 64..512 take the column-per-thread kernel (channelizer8.cu)
        K = 2 * N
        rng = np.random.default_rng(N)
        T = 700
        x = (rng.standard_normal(T * K) + 1j * rng.standard_normal(T * K)).astype(np.complex64)
        L = orc.lib()
        q = L.firpfbch_crcf_create_kaiser(0, K, 7, 60.0)
        # lib/multichannelrx.cc:98: -0.5f*(float)(N-1) / (float)N in float, times M_PI in double, stored as float
        off = np.float32(float(np.float32(np.float32(-0.5) * np.float32(N - 1)) / np.float32(N)) * np.pi)
        u = L.orc_nco_constrain(off)
        n = np.arange(T * K, dtype=np.uint64)
        th = ((n * np.uint64(u)) & np.uint64(0xffffffff)).astype(np.uint32).astype(np.int32)
        t = (th.astype(np.float64) * (np.pi / 2147483648.0)).astype(np.float32)
        xm = (x * (np.cos(t.astype(np.float64)) - 1j * np.sin(t.astype(np.float64)))).astype(np.complex64)
        ref = np.zeros((T, K), np.complex64)
        y = np.zeros(K, np.complex64)
        for b in range(T):
            xi = np.ascontiguousarray(xm[b * K:(b + 1) * K])
            L.firpfbch_crcf_analyzer_execute(q, xi.ctypes.data, y.ctypes.data)
            ref[b] = y
        L.firpfbch_crcf_destroy(q)
        g = pkg.MultichannelRx(N, 64, 16, 4)
        g.execute(x)
        cz = g.read_channelizer()
        g.close()
        assert cz.shape == (N, T)
        err = np.abs(cz - ref[:, :N].T).max() / np.abs(ref).max()
        assert err < 2e-6, (N, err)
        # the same stream in ragged pieces (history + partial-block carry across calls), device input
        import torch
        g = pkg.MultichannelRx(N, 64, 16, 4)
        cuts = [0, 3 * K + 5, 3 * K + 6, 200 * K, 300 * K, 300 * K + K // 2, T * K]     # [200K, 300K) takes the zero-copy path
        got = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if (a % 2) == 0:
                d = torch.from_numpy(x[a:b].view(np.float32).copy()).cuda()
                g.execute_device(d.data_ptr(), b - a)
            else:
                g.execute(x[a:b])
            cz2 = g.read_channelizer()
            if cz2.size:
                got.append(cz2)
        g.close()
        cz2 = np.concatenate(got, axis=1)
        assert cz2.shape == (N, T)
        assert np.array_equal(cz2, cz), N


def test_batched_single_link_sync_matches_multichannel_oracle_streams():
    """b2_ofdmsync_*: each stream is an independent ofdmflexframesync; feed it the oracle
    channelizer's per-channel streams and compare with the oracle's frames"""
    from b2 import pkg
    case = CASES["c2_8ch_h128"]
    N, M, cp, taper = case[:4]
    x = make_input(case)
    fo, po, _ = run_oracle(case, x)
    g = pkg.MultichannelRx(N, M, cp, taper)
    g.execute(x)
    g.poll()
    cz = g.read_channelizer()
    g.close()
    s = pkg.OfdmSync(M, cp, taper, streams=N)
    s.execute(cz)
    fg, pg = s.poll()
    s.close()
    assert_frames_equal(fo, po, fg, pg)


def test_error_codes_match_reference_throws():
    from b2 import pkg
    for args in ((0, 64, 16, 4), (2, 6, 2, 0), (2, 64, 0, 0), (2, 64, 4, 8)):
        with pytest.raises(pkg.B2Error) as e:
            pkg.MultichannelRx(*args)
        assert e.value.code == -1
    with pytest.raises(pkg.B2Error) as e:
        pkg.MultichannelRx(43, 64, 16, 4)         # legal for liquid, outside the CUDA path (prime factor 43 > 41)
    assert e.value.code == -2
    with pytest.raises(pkg.B2Error) as e:
        pkg.MultichannelRx(2, 86, 6, 4)           # M = 2 * 43
    assert e.value.code == -2
    pkg.MultichannelRx(3, 48, 6, 4).close()       # odd channel counts and the reference's default M are fine


def test_stagewise_and_sharded_world1_equal_monolithic():
    """b2_mcrx_channelize_device + b2_mcrx_sync_device, and the ShardedMultichannelRx plumbing with
    a single rank, reproduce b2_mcrx_execute"""
    import importlib
    import torch
    from b2 import pkg
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    case = CASES["c2_8ch_h128"]
    N, M, cp, taper = case[:4]
    K = 2 * N
    x = make_input(case)
    x = x[:(len(x) // (2 * K)) * 2 * K]
    fo, po, _ = run_oracle(case, x)
    T = len(x) // K
    H = sh.HALO_BLOCKS
    # two calls of T/2 blocks each through the sharded front end
    rxs = sh.ShardedMultichannelRx(N, M, cp, taper, T // 2, 0, 1, device=0)
    pad = np.concatenate([np.zeros(H * K, np.complex64), x])
    for call in range(2):
        seg = pad[call * (T // 2) * K:(call * (T // 2) + H + T // 2) * K]
        d = torch.from_numpy(seg.copy()).cuda()
        rxs.execute_device(d)
    fg, pg = rxs.poll()
    rxs.close()
    assert_frames_equal(fo, po, fg, pg)
    # stage-wise calls on one handle
    g = pkg.MultichannelRx(N, M, cp, taper)
    d = torch.from_numpy(pad.copy()).cuda()
    out = torch.empty((N, T), dtype=torch.complex64, device="cuda")
    g.channelize_device(d.data_ptr(), T, -H * K, out.data_ptr(), T)
    torch.cuda.synchronize()
    g.sync_device(out.data_ptr(), T, T)
    fg, pg = g.poll()
    g.close()
    assert_frames_equal(fo, po, fg, pg)


def test_frame_without_payload_symbols_on_every_synchroniser():
    """payload_len 0 with check NONE and no FEC: a legal header announcing ZERO payload symbols (the frame generator of
    this library emits it; liquid's synchroniser would wait for ever).  Every synchroniser kernel -- frame-parallel,
    serial-chain, generic -- must complete the frame with the first OFDM symbol after the header, not hang, and carry on
    with the next frame."""
    from b2 import pkg
    LIQUID_CRC_NONE, LIQUID_CRC_32 = 1, 6
    for M, cp, taper in ((512, 64, 16), (64, 16, 4)):
        gen = pkg.OfdmGen(M, cp, taper)
        parts = [np.zeros(700, np.complex64)]
        for k, (plen, check) in enumerate(((0, LIQUID_CRC_NONE), (40, LIQUID_CRC_32), (0, LIQUID_CRC_NONE))):
            hdr = np.arange(8, dtype=np.uint8) + k
            nsym = gen.assemble(hdr, np.arange(plen, dtype=np.uint8), check, FEC_NONE, FEC_NONE, MOD_QPSK)
            x, last = gen.write(nsym + 2)
            parts += [x * 0.5, np.zeros(3 * (M + cp) + 17 * k, np.complex64)]
        gen.close()
        x = np.concatenate(parts)
        results = []
        for env in ({}, {"B2_SYNC_LEGACY": "1"}, {"B2_SYNC_GENERIC": "1"}):
            os.environ.update(env)
            try:
                rx = pkg.OfdmSync(M, cp, taper, streams=1)
                for c in (x[:len(x) // 3], x[len(x) // 3:]):
                    rx.execute(c)
                fr, pl = rx.poll()
                rx.close()
            finally:
                for k_ in env:
                    os.environ.pop(k_, None)
            assert len(fr) == 3, (M, env, len(fr))
            assert list(fr["header_valid"]) == [1, 1, 1] and list(fr["payload_len"]) == [0, 40, 0]
            assert int(fr["payload_valid"][1]) == 1 and np.array_equal(pl[:40], np.arange(40, dtype=np.uint8))
            assert [int(h[0]) for h in fr["header"]] == [0, 1, 2]
            results.append(fr)
        for other in results[1:]:
            for name in ("detect_index", "complete_index", "payload_len", "header"):
                assert np.array_equal(results[0][name], other[name]), (M, name)


def sharded_chunks(x, K, tc, steps, rank, world, halo):
    """device tensors [halo | chunk] of the chunks rank, rank + world, ... of stream x (zeros before the stream)"""
    import torch
    pad = np.concatenate([np.zeros(halo * K, np.complex64), x])
    out = []
    for i in range(steps):
        g = i * world + rank
        out.append(torch.from_numpy(pad[g * tc * K:(g * tc + halo + tc) * K].copy()).cuda())
    return out


def test_sharded_rx_single_rank_equals_oracle():
    """the pipelined multi-GPU receiver (b2_mcrx_shard_*: round-robin chunks, channelizer storing into the exchange
    slots, synchronisers over the slots) with one rank, over two calls: same records as the oracle"""
    import importlib
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    for name in ("c5_shape_32ch_qam64", "c2_8ch_h128"):
        case = CASES[name]
        N, M, cp, taper = case[:4]
        K = 2 * N
        x = make_input(case)
        steps, calls = 3, 2
        tc = (len(x) // K) // (steps * calls)
        x = x[:tc * steps * calls * K]
        fo, po, _ = run_oracle(case, x)
        rx = sh.ShardedRx(N, M, cp, taper, tc, steps, 0, 1, device=0)
        bufs = sharded_chunks(x, K, tc, steps * calls, 0, 1, sh.HALO_BLOCKS)
        frames, pls = [], []
        for c in range(calls):
            rx.execute_device([b.data_ptr() for b in bufs[c * steps:(c + 1) * steps]])
            f, pl = rx.poll()
            f = f.copy()
            f["payload_offset"] += sum(len(q) for q in pls)
            frames.append(f); pls.append(pl.copy())
        rx.close()
        assert_frames_equal(fo, po, np.concatenate(frames), np.concatenate(pls))


def test_sharded_rx_two_gpus():
    """two ranks under torchrun (tests/mgpu_sharded.py): the gathered records equal a single-GPU receiver's"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(root, "tests", "mgpu_sharded.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_full_size_north_star_shape_properties():
    """BASELINE configs[4] shape at full width (256 channels x 512 subcarriers, 64-QAM), too big for
    the oracle receiver to be run in a test, checked through size-independent properties: every
    transmitted frame is decoded with its exact payload, frames of a channel are one frame period
    apart, and the result does not depend on how the call is split (pipeline chunks, ragged calls)."""
    import torch
    import bench
    from b2 import pkg
    w = bench.WORKLOAD
    period, expected, flen = bench.make_period()
    reps = 5
    x = np.tile(period, reps)
    d = torch.from_numpy(x.view(np.float32)).cuda()
    results = []
    for mode in ("one_call", "ragged"):
        g = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], max_batch=len(x))
        if mode == "one_call":
            g.execute_device(d.data_ptr(), len(x))
        else:
            cuts = [0, 2 * len(period) + 1234, 2 * len(period) + 1236, 3 * len(period) + 7 * 512, len(x)]
            for a, b in zip(cuts[:-1], cuts[1:]):
                g.execute(x[a:b])
        fr, pl = g.poll()
        g.close()
        results.append((fr, pl))
    fr, pl = results[0]
    assert len(fr) >= w["N"] * (reps - 1)
    assert int(fr["header_valid"].min()) == 1 and int(fr["payload_valid"].min()) == 1
    for c in range(w["N"]):
        rows = fr[fr["channel"] == c]
        assert len(rows) >= reps - 1
        # the very first detection happens on the start-up seek grid, the later ones on the grid the
        # previous frame's end defines: periodic from the second frame on
        assert np.all(np.diff(rows["detect_index"].astype(np.int64))[1:] == flen), c
        assert np.all(np.diff(rows["complete_index"].astype(np.int64)) == flen), c
        for r in rows:
            o = int(r["payload_offset"])
            assert np.array_equal(pl[o:o + w["payload"]], expected[c][1]), c
    fr2, pl2 = results[1]
    assert len(fr2) == len(fr)
    for name in EXACT:
        assert np.array_equal(fr[name], fr2[name]), name
    assert np.array_equal(pl, pl2)


def test_resampler_ahead_of_the_receiver():
    """msresamp_crcf in front of multichannelrx (the ratio src/multichannel_rx.cc:137-138
    computes).  A capture taken at 1.07x the receiver's rate is brought back by b2_mcrx_set_resampler; the
    frames must equal the oracle receiver run on the resampled stream, whatever the call sizes, from host or
    device memory, and every frame must survive the two rate changes."""
    import torch
    import orc
    from b2 import pkg
    for name in ("c2_8ch_h128", "c5_shape_32ch_qam64"):
        case = CASES[name]
        N, M, cp, taper = case[:4]
        x = make_input(case)
        xr = orc.msresamp(x, np.float32(1.07))                    # what a radio at 1.07x would have captured
        rate = np.float32(1.0 / 1.07)
        rs = pkg.MsResamp(rate)
        y = rs.execute(xr)
        rs.close()
        fo, po, _ = run_oracle(case, y)
        assert len(fo) == N * case[8] and int(fo["payload_valid"].sum()) == len(fo)
        rng = np.random.default_rng(12)
        for mode in ("one_call", "ragged", "device", "small_batch"):
            g = pkg.MultichannelRx(N, M, cp, taper, max_batch=5000 if mode == "small_batch" else 0)
            g.set_resampler(rate)
            if mode == "ragged":
                i = 0
                while i < len(xr):
                    c = int(rng.integers(1, 20000))
                    g.execute(xr[i:i + c])
                    i += c
            elif mode == "device":
                d = torch.from_numpy(xr.view(np.float32)).cuda()
                g.execute_device(d.data_ptr(), len(xr))
            else:
                g.execute(xr)
            fg, pg = g.poll()
            g.set_resampler(0.0)                                   # removing the stage leaves a plain receiver
            g.close()
            assert_frames_equal(fo, po, fg, pg)


def test_two_devices_in_one_process():
    """handles on different devices of one process (the C ABI takes a device ordinal): per-device kernel
    attributes, SM partitions and streams must not leak from one device to the other"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from b2 import pkg
    case = CASES["c5_shape_32ch_qam64"]
    N, M, cp, taper = case[:4]
    x = make_input(case)
    fo, po, _ = run_oracle(case, x)
    rxs = [pkg.MultichannelRx(N, M, cp, taper, device=d) for d in (1, 0)]
    half = len(x) // 2 + 7
    for rx in rxs:
        rx.execute(x[:half])
    for rx in reversed(rxs):
        rx.execute(x[half:])
    for rx in rxs:
        fg, pg = rx.poll()
        assert_frames_equal(fo, po, fg, pg)
        rx.close()
    # the transmit side and the resampler on the second device
    rs0, rs1 = pkg.MsResamp(np.float32(1.07), device=0), pkg.MsResamp(np.float32(1.07), device=1)
    assert np.array_equal(rs0.execute(x[:100000]), rs1.execute(x[:100000]))
    rs0.close(); rs1.close()


def test_worker_pairs_under_corruption():
    """speculative workers of the synchronisers (frame-parallel kernel: workers starting at predicted frame boundaries;
    serial-chain kernel: worker pairs on alternate frames): long runs of back-to-back frames with OFDM symbols wiped at
    random places -- lost preambles, invalid headers, failed CRCs, i.e. predictions that do not hold and stretches the
    stitcher has to redo serially -- plus noise, fed in ragged calls.  Records must equal the oracle's in every
    configuration."""
    from b2 import pkg
    rng = np.random.default_rng(21)
    for case in ((8, 256, 32, 8, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150, 9, 0.02),
                 (4, 512, 64, 16, MOD_QAM16, FEC_NONE, FEC_NONE, 400, 8, 0.01)):
        N, M, cp, taper = case[:4]
        K, W = 2 * N, M + cp
        import orc
        from test_oracle_loopback import frame_symbols
        bps = {MOD_QPSK: 2, MOD_QAM16: 4}[case[4]]
        nsym = frame_symbols(M, bps, orc.lib().orc_packetizer_enc_len(case[7], 6, case[5], case[6]))
        x = make_input(case).copy()
        # frame f of every channel starts near channel sample f * nsym * W: wipe parts of a header symbol, of a
        # payload symbol, of a preamble symbol and of a last payload symbol (all channels at once)
        for f, sym, frac in ((1, 3, 0.6), (3, 5, 0.5), (4, 1, 0.7), (6, 3, 0.9), (7, nsym - 1, 0.5)):
            a = ((f * nsym + sym) * W + 13) * K
            x[a:a + int(frac * W * K)] = 0
        fo, po, _ = run_oracle(case, x)
        assert len(fo) >= N * 3
        assert int((fo["header_valid"] == 0).sum()) >= N and int((fo["payload_valid"] == 0).sum()) >= N
        chunks, left = [], len(x)
        while left:
            c = int(min(left, rng.integers(1, 40000)))
            chunks.append(c)
            left -= c
        # frame-parallel kernel with its default number of workers per channel, with 3, and serial (1); then the
        # serial-chain kernel of round 1 with and without its worker pairs
        for env in ({}, {"B2_SYNC_K": "3"}, {"B2_SYNC_K": "1"}, {"B2_SYNC_LEGACY": "1", "B2_SYNC_WORKERS": "2"}, {"B2_SYNC_LEGACY": "1", "B2_SYNC_WORKERS": "1"}):
            os.environ.update(env)
            try:
                fg, pg, _ = run_gpu(case, x)
                assert_frames_equal(fo, po, fg, pg)
                fg, pg, _ = run_gpu(case, x, chunks)
                assert_frames_equal(fo, po, fg, pg)
            finally:
                for k_ in env:
                    os.environ.pop(k_, None)
