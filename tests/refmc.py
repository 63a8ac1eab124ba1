"""ctypes view of tests/shim/mc_shim.cc (class-API driver).  Loaded over
oracle/_ref/libref_mc.so it runs the reference's own lib/multichannel{rx,tx}.cc on the CPU
oracle; loaded over the product's shim library it runs the CUDA-backed classes."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# liquid enum values used in tests (include/liquid/liquid.h)
CRC_32 = 6
FEC_NONE, FEC_HAMMING128, FEC_GOLAY2412, FEC_CONV_V27 = 1, 6, 7, 11
MOD_QAM16, MOD_QAM64, MOD_QAM256 = 27, 29, 31
MOD_BPSK, MOD_QPSK = 39, 40


class ShimFrame(C.Structure):
    _fields_ = [("channel", C.c_uint32), ("header_valid", C.c_int32), ("payload_valid", C.c_int32),
                ("payload_len", C.c_uint32), ("header", C.c_uint8 * 8),
                ("evm", C.c_float), ("rssi", C.c_float), ("cfo", C.c_float),
                ("mod_scheme", C.c_uint32), ("mod_bps", C.c_uint32), ("check", C.c_uint32),
                ("fec0", C.c_uint32), ("fec1", C.c_uint32),
                ("detect_index", C.c_uint64), ("complete_index", C.c_uint64),
                ("payload_offset", C.c_uint64)]


FRAME_DTYPE = np.dtype([("channel", "<u4"), ("header_valid", "<i4"), ("payload_valid", "<i4"),
                        ("payload_len", "<u4"), ("header", "u1", (8,)),
                        ("evm", "<f4"), ("rssi", "<f4"), ("cfo", "<f4"),
                        ("mod_scheme", "<u4"), ("mod_bps", "<u4"), ("check", "<u4"),
                        ("fec0", "<u4"), ("fec1", "<u4"),
                        ("detect_index", "<u8"), ("complete_index", "<u8"),
                        ("payload_offset", "<u8")], align=True)
assert FRAME_DTYPE.itemsize == C.sizeof(ShimFrame)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class McLib:
    def __init__(self, path):
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.mcshim_rx_create.restype = C.c_void_p
        L.mcshim_rx_create.argtypes = [C.c_uint] * 4
        L.mcshim_rx_destroy.argtypes = [C.c_void_p]
        L.mcshim_rx_reset.argtypes = [C.c_void_p]
        L.mcshim_rx_execute.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_uint64]
        L.mcshim_rx_num_frames.restype = C.c_uint64
        L.mcshim_rx_num_frames.argtypes = [C.c_void_p]
        L.mcshim_rx_payload_bytes.restype = C.c_uint64
        L.mcshim_rx_payload_bytes.argtypes = [C.c_void_p]
        L.mcshim_rx_get_frames.argtypes = [C.c_void_p, C.c_void_p]
        L.mcshim_rx_get_payloads.argtypes = [C.c_void_p, C.c_void_p]
        L.mcshim_rx_clear.argtypes = [C.c_void_p]
        L.mcshim_rx_enable_symbol_tap.argtypes = [C.c_void_p, C.c_int]
        L.mcshim_rx_num_symbols.restype = C.c_uint64
        L.mcshim_rx_num_symbols.argtypes = [C.c_void_p]
        L.mcshim_rx_get_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcshim_tx_create.restype = C.c_void_p
        L.mcshim_tx_create.argtypes = [C.c_uint] * 4
        L.mcshim_tx_destroy.argtypes = [C.c_void_p]
        L.mcshim_tx_reset.argtypes = [C.c_void_p]
        L.mcshim_tx_is_ready.argtypes = [C.c_void_p, C.c_uint]
        L.mcshim_tx_update.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.mcshim_tx_generate.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_uint64]
        L.mcshim_tx_run.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_uint, C.c_int, C.c_int, C.c_int,
                                    C.c_uint64, C.c_uint64, C.c_uint, C.c_float]
        L.mcshim_frame_data.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint]

    def frame_data(self, seed, channel, pid, payload_len):
        h = np.zeros(8, np.uint8)
        p = np.zeros(max(payload_len, 1), np.uint8)
        self.lib.mcshim_frame_data(seed, channel, pid, h.ctypes.data, p.ctypes.data, payload_len)
        return h, p[:payload_len]


class McTx:
    def __init__(self, mclib, N, M, cp, taper):
        self.L = mclib.lib
        self.N, self.M, self.cp, self.taper = N, M, cp, taper
        self.h = self.L.mcshim_tx_create(N, M, cp, taper)
        if not self.h:
            raise ValueError("multichanneltx constructor threw")

    def close(self):
        if self.h:
            self.L.mcshim_tx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def reset(self):
        self.L.mcshim_tx_reset(self.h)

    def is_ready(self, c):
        return self.L.mcshim_tx_is_ready(self.h, c)

    def update(self, c, header, payload, mod, fec0, fec1):
        header = np.ascontiguousarray(header, np.uint8)
        payload = np.ascontiguousarray(payload, np.uint8)
        return self.L.mcshim_tx_update(self.h, c, header.ctypes.data, payload.ctypes.data, len(payload), mod, fec0, fec1)

    def generate(self, ncalls):
        out = np.zeros(ncalls * 2 * self.N, np.complex64)
        self.L.mcshim_tx_generate(self.h, _fptr(out), ncalls)
        return out

    def run(self, ncalls, payload_len, mod, fec0, fec1, seed=0xB2000000, channel_mask=(1 << 64) - 1, max_frames=0, gain=1.0):
        out = np.zeros(ncalls * 2 * self.N, np.complex64)
        self.L.mcshim_tx_run(self.h, _fptr(out), ncalls, payload_len, mod, fec0, fec1, seed, channel_mask, max_frames, gain)
        return out


class McRx:
    def __init__(self, mclib, N, M, cp, taper):
        self.L = mclib.lib
        self.N, self.M, self.cp, self.taper = N, M, cp, taper
        self.h = self.L.mcshim_rx_create(N, M, cp, taper)
        if not self.h:
            raise ValueError("multichannelrx constructor threw")

    def close(self):
        if self.h:
            self.L.mcshim_rx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def reset(self):
        self.L.mcshim_rx_reset(self.h)

    def tap_symbols(self, enable=True):
        return self.L.mcshim_rx_enable_symbol_tap(self.h, int(enable))

    def execute(self, x, chunk=0):
        x = np.ascontiguousarray(x, np.complex64)
        self.L.mcshim_rx_execute(self.h, _fptr(x), len(x), chunk)

    def frames(self, clear=True):
        n = self.L.mcshim_rx_num_frames(self.h)
        fr = np.zeros(n, FRAME_DTYPE)
        pl = np.zeros(self.L.mcshim_rx_payload_bytes(self.h), np.uint8)
        if n:
            self.L.mcshim_rx_get_frames(self.h, fr.ctypes.data)
        if len(pl):
            self.L.mcshim_rx_get_payloads(self.h, pl.ctypes.data)
        if clear:
            self.L.mcshim_rx_clear(self.h)
        return fr, pl

    def symbols(self):
        n = self.L.mcshim_rx_num_symbols(self.h)
        ch = np.zeros(n, np.uint32)
        idx = np.zeros(n, np.uint64)
        X = np.zeros((n, self.M), np.complex64)
        if n:
            self.L.mcshim_rx_get_symbols(self.h, ch.ctypes.data, idx.ctypes.data, X.ctypes.data)
        return ch, idx, X


def payload_of(fr, pl, i):
    o = int(fr["payload_offset"][i])
    return pl[o:o + int(fr["payload_len"][i])]


_ref = None


def ref_lib():
    """the reference's lib/*.cc over the oracle (oracle/_ref/libref_mc.so)"""
    global _ref
    if _ref is None:
        C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"), mode=C.RTLD_LOCAL)
        _ref = McLib(os.path.join(ROOT, "oracle", "_ref", "libref_mc.so"))
    return _ref
