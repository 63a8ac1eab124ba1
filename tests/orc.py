"""ctypes bindings of the CPU oracle's C functions (oracle/liborc.so) -- test infrastructure."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None


class Packetizer(C.Structure):
    _fields_ = [("msg_len", C.c_uint), ("packet_len", C.c_uint), ("check", C.c_uint), ("crc_len", C.c_uint),
                ("fs", C.c_uint * 2), ("dec_len", C.c_uint * 2), ("enc_len", C.c_uint * 2), ("depth", C.c_uint * 2),
                ("buf0", C.c_void_p), ("buf1", C.c_void_p)]


class Modem(C.Structure):
    _fields_ = [("scheme", C.c_uint), ("bps", C.c_uint), ("M", C.c_uint), ("m_i", C.c_uint), ("m_q", C.c_uint),
                ("alpha", C.c_float), ("ref", C.c_float * 8), ("r", C.c_float * 2), ("x_hat", C.c_float * 2)]


class Mseq(C.Structure):
    _fields_ = [(n, C.c_uint) for n in ("m", "g", "a", "n", "v", "b")]


def lib():
    global _L
    if _L is not None:
        return _L
    L = C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"), mode=C.RTLD_LOCAL)
    L.orc_crc32.restype = C.c_uint32
    L.orc_crc32.argtypes = [C.c_void_p, C.c_uint]
    for f in ("orc_hamming128_encode_symbol", "orc_hamming128_decode_symbol",
              "orc_golay2412_encode_symbol", "orc_golay2412_decode_symbol"):
        getattr(L, f).restype = C.c_uint
        getattr(L, f).argtypes = [C.c_uint]
    L.orc_fec_enc_len.restype = C.c_uint
    L.orc_fec_enc_len.argtypes = [C.c_uint, C.c_uint]
    L.orc_fec_encode.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    L.orc_fec_decode.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    L.orc_interleave.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int]
    L.orc_scramble.argtypes = [C.c_void_p, C.c_uint]
    L.orc_packetizer_enc_len.restype = C.c_uint
    L.orc_packetizer_enc_len.argtypes = [C.c_uint] * 4
    L.orc_packetizer_init.argtypes = [C.POINTER(Packetizer)] + [C.c_uint] * 4
    L.orc_packetizer_free.argtypes = [C.POINTER(Packetizer)]
    L.orc_packetizer_encode.argtypes = [C.POINTER(Packetizer), C.c_void_p, C.c_void_p]
    L.orc_packetizer_decode.restype = C.c_int
    L.orc_packetizer_decode.argtypes = [C.POINTER(Packetizer), C.c_void_p, C.c_void_p]
    L.orc_firdes_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.orc_kaiser_beta_As.restype = C.c_float
    L.orc_kaiser_beta_As.argtypes = [C.c_float]
    L.orc_fft_create.restype = C.c_void_p
    L.orc_fft_create.argtypes = [C.c_uint, C.c_int]
    L.orc_fft_destroy.argtypes = [C.c_void_p]
    L.orc_fft_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_modem_init.argtypes = [C.POINTER(Modem), C.c_uint]
    L.orc_modem_modulate.argtypes = [C.POINTER(Modem), C.c_uint]
    L.orc_modem_demodulate.restype = C.c_uint
    L.orc_mseq_init_default.argtypes = [C.POINTER(Mseq), C.c_uint]
    L.orc_mseq_advance.restype = C.c_uint
    L.orc_mseq_advance.argtypes = [C.POINTER(Mseq)]
    L.orc_nco_constrain.restype = C.c_uint32
    L.orc_nco_constrain.argtypes = [C.c_float]
    L.ofdmframe_init_default_sctype.argtypes = [C.c_uint, C.c_void_p]
    L.firpfbch_crcf_create_kaiser.restype = C.c_void_p
    L.firpfbch_crcf_create_kaiser.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float]
    L.firpfbch_crcf_destroy.argtypes = [C.c_void_p]
    L.firpfbch_crcf_analyzer_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.firpfbch_crcf_synthesizer_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.msresamp_crcf_create.restype = C.c_void_p
    L.msresamp_crcf_create.argtypes = [C.c_float, C.c_float]
    L.msresamp_crcf_destroy.argtypes = [C.c_void_p]
    L.msresamp_crcf_reset.argtypes = [C.c_void_p]
    L.msresamp_crcf_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]
    L.orc_msresamp_get_design.restype = C.c_uint
    L.orc_msresamp_get_design.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint64)]
    # single-link framer
    L.ofdmflexframegen_create.restype = C.c_void_p
    L.ofdmflexframegen_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p]
    L.ofdmflexframegen_destroy.argtypes = [C.c_void_p]
    L.ofdmflexframegen_setprops.argtypes = [C.c_void_p, C.c_void_p]
    L.ofdmflexframegen_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    L.ofdmflexframegen_write.restype = C.c_int
    L.ofdmflexframegen_write.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.ofdmflexframegen_writesymbol.restype = C.c_int
    L.ofdmflexframegen_writesymbol.argtypes = [C.c_void_p, C.c_void_p]
    L.ofdmflexframegen_getframelen.restype = C.c_uint
    L.ofdmflexframegen_getframelen.argtypes = [C.c_void_p]
    L.ofdmflexframegen_is_assembled.argtypes = [C.c_void_p]
    _L = L
    return L


def u8(a):
    return np.ascontiguousarray(a, np.uint8)


def crc32(msg):
    m = u8(msg)
    return lib().orc_crc32(m.ctypes.data, len(m))


def fec_encode(scheme, msg):
    m = u8(msg)
    out = np.zeros(lib().orc_fec_enc_len(scheme, len(m)) + 8, np.uint8)
    lib().orc_fec_encode(scheme, len(m), m.ctypes.data, out.ctypes.data)
    return out[:-8]


def fec_decode(scheme, dec_len, enc):
    e = u8(np.concatenate([u8(enc), np.zeros(8, np.uint8)]))
    out = np.zeros(dec_len + 8, np.uint8)
    lib().orc_fec_decode(scheme, dec_len, e.ctypes.data, out.ctypes.data)
    return out[:dec_len]


def interleave(x, depth=4, decode=False):
    y = u8(x).copy()
    lib().orc_interleave(y.ctypes.data, len(y), depth, int(decode))
    return y


def packetizer_encode(msg, check, fec0, fec1):
    m = u8(msg)
    p = Packetizer()
    lib().orc_packetizer_init(C.byref(p), len(m), check, fec0, fec1)
    out = np.zeros(p.packet_len, np.uint8)
    lib().orc_packetizer_encode(C.byref(p), m.ctypes.data, out.ctypes.data)
    lib().orc_packetizer_free(C.byref(p))
    return out


def packetizer_decode(pkt, msg_len, check, fec0, fec1):
    k = u8(pkt)
    p = Packetizer()
    lib().orc_packetizer_init(C.byref(p), msg_len, check, fec0, fec1)
    assert p.packet_len == len(k)
    out = np.zeros(msg_len, np.uint8)
    ok = lib().orc_packetizer_decode(C.byref(p), k.ctypes.data, out.ctypes.data)
    lib().orc_packetizer_free(C.byref(p))
    return out, bool(ok)


def firdes_kaiser(n, fc, As, mu=0.0):
    h = np.zeros(n, np.float32)
    lib().orc_firdes_kaiser(n, fc, As, mu, h.ctypes.data)
    return h


def fft(x, backward=False):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.zeros_like(x)
    q = lib().orc_fft_create(len(x), 1 if backward else -1)
    lib().orc_fft_execute(q, x.ctypes.data, y.ctypes.data)
    lib().orc_fft_destroy(q)
    return y


def default_sctype(M):
    p = np.zeros(M, np.uint8)
    lib().ofdmframe_init_default_sctype(M, p.ctypes.data)
    return p


def msresamp(x, rate, As=60.0):
    x = np.ascontiguousarray(x, np.complex64)
    q = lib().msresamp_crcf_create(rate, As)
    y = np.zeros(int(len(x) * rate) + 64, np.complex64)
    ny = C.c_uint(0)
    lib().msresamp_crcf_execute(q, x.ctypes.data, len(x), y.ctypes.data, C.byref(ny))
    lib().msresamp_crcf_destroy(q)
    return y[:ny.value]
