"""Parity of the CUDA transmit path and resampler against the CPU oracle: frame generator
(ofdmflexframegen), synthesis channelizer + NCO (multichanneltx), msresamp_crcf.

Bars: samples within 1e-5 of the oracle relative to the signal peak; decoding the CUDA
transmitter's output with the ORACLE receiver gives bit-exact frames; full CUDA tx -> CUDA rx
loopback decodes every frame."""
import ctypes as C

import numpy as np
import pytest

import orc
from refmc import (McRx, McTx, ref_lib, payload_of, CRC_32, FEC_NONE, FEC_HAMMING128, FEC_GOLAY2412, FEC_CONV_V27,
                   MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256)

pytestmark = pytest.mark.gpu
SEED = 0xB2000000

TX_CASES = {
    "c2_8ch_h128": (8, 64, 16, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 150),
    "c3_16ch_qam16_v27": (16, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, 300),
    "c5_shape_32ch_qam64": (32, 512, 64, 16, MOD_QAM64, FEC_NONE, FEC_NONE, 1200),
    "1ch_qam256_golay": (1, 128, 16, 4, MOD_QAM256, FEC_GOLAY2412, FEC_NONE, 77),
    # not powers of two: the reference programs' default OFDM shape and odd channel counts
    "ref_defaults_m48_1ch": (1, 48, 6, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 100),
    "n3_m48_qam16": (3, 48, 6, 4, MOD_QAM16, FEC_NONE, FEC_NONE, 120),
    "n6_m80_h128": (6, 80, 10, 4, MOD_QPSK, FEC_NONE, FEC_HAMMING128, 100),
}


def drive(tx_update, tx_ready, tx_generate, N, ncalls, plen, mod, fec0, fec1, max_frames, batch):
    """the src/multichannel_tx.cc:163-216 loop, `batch` GenerateSamples calls at a time at most,
    never across a symbol boundary (so both implementations see identical call sequences)"""
    L = ref_lib()
    pid = [0] * N
    out = []
    done = 0
    while done < ncalls:
        for c in range(N):
            if pid[c] < max_frames and tx_ready(c):
                h, p = L.frame_data(SEED, c, pid[c], plen)
                tx_update(c, h, p, mod, fec0, fec1)
                pid[c] += 1
        n = min(batch, ncalls - done)
        out.append(tx_generate(n))
        done += n
    return np.concatenate(out)


@pytest.mark.parametrize("name", list(TX_CASES))
def test_multichanneltx_matches_oracle_and_oracle_rx_decodes_it(name):
    from b2 import pkg
    N, M, cp, taper, mod, fec0, fec1, plen = TX_CASES[name]
    W = M + cp
    ncalls = W * 30 if M >= 256 else W * 70
    L = ref_lib()
    otx = McTx(L, N, M, cp, taper)
    xo = drive(otx.update, otx.is_ready, otx.generate, N, ncalls, plen, mod, fec0, fec1, 2, W)
    otx.close()
    gtx = pkg.MultichannelTx(N, M, cp, taper)
    xg = drive(gtx.update, gtx.is_ready, gtx.generate, N, ncalls, plen, mod, fec0, fec1, 2, W)
    gtx.close()
    assert len(xg) == len(xo)
    err = np.abs(xg - xo).max() / np.abs(xo).max()
    assert err < 1e-5, err
    # the oracle receiver decodes the CUDA transmitter's samples into the same frames
    rxa = McRx(L, N, M, cp, taper); rxa.execute(xo); fa, pa = rxa.frames(); rxa.close()
    rxb = McRx(L, N, M, cp, taper); rxb.execute(xg); fb, pb = rxb.frames(); rxb.close()
    assert len(fa) == 2 * N
    for k in ("channel", "header_valid", "payload_valid", "header", "detect_index", "complete_index"):
        assert np.array_equal(fa[k], fb[k]), k
    assert np.array_equal(pa, pb) and int(fa["payload_valid"].sum()) == len(fa)


def test_generate_is_chunking_invariant_and_ready_flags_follow_reference():
    from b2 import pkg
    N, M, cp, taper, mod, fec0, fec1, plen = TX_CASES["c2_8ch_h128"]
    W = M + cp
    L = ref_lib()
    otx = McTx(L, N, M, cp, taper)
    gtx = pkg.MultichannelTx(N, M, cp, taper)
    rng = np.random.default_rng(3)
    pid = [0] * N
    xo, xg = [], []
    done = 0
    while done < W * 45:
        for c in range(N):
            ro, rg = otx.is_ready(c), gtx.is_ready(c)
            assert ro == rg, (done, c)
            if ro and pid[c] < 3 and rng.random() < 0.5:          # channels arm at different times
                h, p = L.frame_data(SEED, c, pid[c], plen)
                otx.update(c, h, p, mod, fec0, fec1)
                gtx.update(c, h, p, mod, fec0, fec1)
                pid[c] += 1
        n = int(rng.integers(1, 2 * W))                            # crosses symbol boundaries freely
        xo.append(otx.generate(n))
        xg.append(gtx.generate(n))
        done += n
    xo, xg = np.concatenate(xo), np.concatenate(xg)
    assert np.abs(xg - xo).max() / np.abs(xo).max() < 1e-5
    # update on a busy channel is refused (the reference warns and returns), bad channel throws
    for c in range(N):
        if not gtx.is_ready(c):
            with pytest.raises(pkg.B2Error) as e:
                gtx.update(c, np.zeros(8, np.uint8), np.zeros(4, np.uint8), mod, fec0, fec1)
            assert e.value.code == -5
            break
    with pytest.raises(pkg.B2Error) as e:
        gtx.is_ready(N)
    assert e.value.code == -1
    otx.close(); gtx.close()


def test_full_cuda_loopback_decodes_every_frame():
    from b2 import pkg
    N, M, cp, taper, mod, fec0, fec1, plen = TX_CASES["c3_16ch_qam16_v27"]
    W = M + cp
    L = ref_lib()
    gtx = pkg.MultichannelTx(N, M, cp, taper)
    x = drive(gtx.update, gtx.is_ready, gtx.generate, N, W * 80, plen, mod, fec0, fec1, 3, 4 * W)
    gtx.close()
    grx = pkg.MultichannelRx(N, M, cp, taper)
    grx.execute(x / N)
    fr, pl = grx.poll()
    grx.close()
    assert len(fr) == 3 * N
    assert int(fr["header_valid"].sum()) == len(fr) and int(fr["payload_valid"].sum()) == len(fr)
    for i in range(len(fr)):
        c = int(fr["channel"][i]); pid = int(fr["header"][i][0]) * 256 + int(fr["header"][i][1])
        h, p = L.frame_data(SEED, c, pid, plen)
        assert np.array_equal(fr["header"][i], h) and np.array_equal(payload_of(fr, pl, i), p)


@pytest.mark.parametrize("M,cp,taper,mod,fec0,fec1,plen", [(64, 16, 4, MOD_QPSK, FEC_NONE, FEC_NONE, 100),
                                                           (512, 64, 16, MOD_QAM256, FEC_NONE, FEC_NONE, 1200),
                                                           (256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_HAMMING128, 64)])
def test_single_link_framegen_matches_oracle(M, cp, taper, mod, fec0, fec1, plen):
    from b2 import pkg
    L = orc.lib()
    W = M + cp
    rng = np.random.default_rng(M)
    header = rng.integers(0, 256, 8, dtype=np.uint8)
    payload = rng.integers(0, 256, plen, dtype=np.uint8)
    props = (C.c_uint * 4)(CRC_32, fec0, fec1, mod)
    fg = L.ofdmflexframegen_create(M, cp, taper, None, props)
    L.ofdmflexframegen_assemble(fg, header.ctypes.data, payload.ctypes.data, plen)
    ref = []
    while True:
        buf = np.zeros(W, np.complex64)
        last = L.ofdmflexframegen_write(fg, buf.ctypes.data, W)
        ref.append(buf)
        if last:
            break
    L.ofdmflexframegen_destroy(fg)
    ref = np.concatenate(ref)
    g = pkg.OfdmGen(M, cp, taper)
    nsym = g.assemble(header, payload, CRC_32, fec0, fec1, mod)
    assert nsym * W == len(ref) and g.is_assembled()
    a, last_a = g.write(3)
    b, last_b = g.write(nsym - 3)
    out = np.concatenate([a, b])
    assert last_a == 0 and last_b == 1 and not g.is_assembled()
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-5
    z, _ = g.write(2)                                     # nothing assembled: silence
    assert not z.any()
    g.close()


def test_msresamp_matches_oracle_and_chunks():
    from b2 import pkg
    rng = np.random.default_rng(9)
    n = 50000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for rate in (1.07, 0.93, 1.9, 0.5):
        ref = orc.msresamp(x, np.float32(rate))
        g = pkg.MsResamp(np.float32(rate))
        y = g.execute(x)
        assert len(y) == len(ref), (rate, len(y), len(ref))
        assert np.abs(y - ref).max() / np.abs(ref).max() < 2e-6
        g.reset()
        parts, i = [], 0
        while i < n:
            c = int(rng.integers(1, 3000))
            parts.append(g.execute(x[i:i + c]))
            i += c
        y2 = np.concatenate(parts)
        assert len(y2) == len(ref) and np.abs(y2 - ref).max() / np.abs(ref).max() < 2e-6
        g.close()
    with pytest.raises(pkg.B2Error) as e:
        pkg.MsResamp(np.float32(3.0))
    assert e.value.code == -2


def test_config4_single_link_through_the_resampler():
    """BASELINE configs[3] composed on the GPU: ofdmflexframegen (M = 512, cp 64, 256-QAM, no FEC) ->
    msresamp_crcf(1.07) -> msresamp_crcf(1/1.07) -> ofdmflexframesync, the way src/flexframe_tx.cc:170,237 /
    src/flexframe_rx.cc:179,240 put the resampler either side of a framer.  Every frame must come back
    with its exact payload, and the resampled waveform must match the oracle resampler."""
    from b2 import pkg
    M, cp, taper, plen, nframes = 512, 64, 16, 1200, 24
    rng = np.random.default_rng(4)
    g = pkg.OfdmGen(M, cp, taper)
    sent, wave = [], []
    for f in range(nframes):
        header = rng.integers(0, 256, 8, dtype=np.uint8)
        payload = rng.integers(0, 256, plen, dtype=np.uint8)
        nsym = g.assemble(header, payload, CRC_32, FEC_NONE, FEC_NONE, MOD_QAM256)
        out, last = g.write(nsym)
        assert last == 1
        wave.append(out)
        wave.append(np.zeros(3 * (M + cp), np.complex64))          # idle gap between packets
        sent.append((header, payload))
    g.close()
    x = np.concatenate(wave)
    up = pkg.MsResamp(np.float32(1.07))
    y = up.execute(x)
    up.close()
    ref = orc.msresamp(x, np.float32(1.07))
    assert len(y) == len(ref) and np.abs(y - ref).max() / np.abs(ref).max() < 2e-6
    down = pkg.MsResamp(np.float32(1.0 / 1.07))
    z = down.execute(y)
    down.close()
    rx = pkg.OfdmSync(M, cp, taper, streams=1, max_batch=len(z))
    rx.execute(z.reshape(1, -1))
    fr, pl = rx.poll()
    rx.close()
    assert len(fr) == nframes, len(fr)
    assert int(fr["header_valid"].min()) == 1 and int(fr["payload_valid"].min()) == 1
    for i, (header, payload) in enumerate(sent):
        assert np.array_equal(fr["header"][i], header)
        o = int(fr["payload_offset"][i])
        assert np.array_equal(pl[o:o + plen], payload)


def test_msresamp_device_paths_equal_host_path():
    """one launch over device memory (any length, history carried on the device) == the staged host path"""
    import torch
    from b2 import pkg
    rng = np.random.default_rng(10)
    n = 300001
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for rate in (1.07, 0.5, 2.0):
        g = pkg.MsResamp(np.float32(rate))
        ref = g.execute(x)
        d_x = torch.from_numpy(x.view(np.float32)).cuda()
        cap = int(n * rate * 1.01) + 64
        d_y = torch.zeros(2 * cap, dtype=torch.float32, device="cuda")
        for cuts in ([n], [1, 5, 13, 14, 1000, 77777]):
            g.reset()
            d_y.zero_()
            i = o = 0
            for c in cuts + [n - sum(cuts)] if sum(cuts) < n else cuts:
                o += g.execute_device(d_x.data_ptr() + 8 * i, c, d_y.data_ptr() + 8 * o, cap - o)
                i += c
            assert i == n and o == len(ref), (rate, i, o, len(ref))
            assert np.array_equal(d_y[:2 * o].cpu().numpy().view(np.complex64), ref), (rate, cuts)
        g.reset()
        o = g.execute_to_device(x[:1234], d_y.data_ptr(), cap)
        o += g.execute_to_device(x[1234:], d_y.data_ptr() + 8 * o, cap - o)
        assert o == len(ref) and np.array_equal(d_y[:2 * o].cpu().numpy().view(np.complex64), ref)
        g.close()
