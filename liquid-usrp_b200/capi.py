"""ctypes bindings of include/b200_ofdm.h (libb200ofdm.so)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(HERE, "libb200ofdm.so")


class B2Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("b200ofdm error %d: %s" % (code, text))
        self.code = code


# numpy view of b2_frame_rec
FRAME_DTYPE = np.dtype([("channel", "<u4"), ("header_valid", "<i4"), ("payload_valid", "<i4"),
                        ("payload_len", "<u4"), ("header", "u1", (8,)),
                        ("evm", "<f4"), ("rssi", "<f4"), ("cfo", "<f4"),
                        ("mod_scheme", "<u4"), ("mod_bps", "<u4"), ("check", "<u4"),
                        ("fec0", "<u4"), ("fec1", "<u4"),
                        ("detect_index", "<u8"), ("complete_index", "<u8"),
                        ("payload_offset", "<u8")], align=True)

_sz = C.c_size_t
_vp = C.c_void_p

_PROTOS = {
    "b2_last_error": (C.c_char_p, []),
    "b2_version": (C.c_char_p, []),
    "b2_device_count": (C.c_int, []),
    "b2_pinned_alloc": (_vp, [_sz]),
    "b2_pinned_free": (None, [_vp]),
    "b2_mcrx_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, C.c_uint, _vp, C.c_int, _sz, C.POINTER(_vp)]),
    "b2_mcrx_destroy": (C.c_int, [_vp]),
    "b2_mcrx_reset": (C.c_int, [_vp]),
    "b2_mcrx_execute": (C.c_int, [_vp, _vp, _sz]),
    "b2_mcrx_execute_device": (C.c_int, [_vp, _vp, _sz]),
    "b2_mcrx_poll": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz)]),
    "b2_mcrx_poll_view": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz)]),
    "b2_mcrx_tap_symbols": (C.c_int, [_vp, C.c_int, _sz]),
    "b2_mcrx_read_symbols": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "b2_mcrx_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float * 4)]),
    "b2_mcrx_last_launches": (C.c_int, [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "b2_mcrx_read_channelizer": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "b2_mcrx_stream": (_vp, [_vp]),
    "b2_mcrx_channelize_device": (C.c_int, [_vp, _vp, _sz, C.c_int64, _vp, _sz]),
    "b2_mcrx_sync_device": (C.c_int, [_vp, _vp, _sz, _sz]),
    "b2_ofdmsync_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, _vp, C.c_uint, C.c_int, _sz, C.POINTER(_vp)]),
    "b2_ofdmsync_destroy": (C.c_int, [_vp]),
    "b2_ofdmsync_reset": (C.c_int, [_vp]),
    "b2_ofdmsync_execute": (C.c_int, [_vp, _vp, _sz]),
    "b2_ofdmsync_execute_device": (C.c_int, [_vp, _vp, _sz, _sz]),
    "b2_ofdmsync_poll": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz)]),
    "b2_ofdmsync_poll_view": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz)]),
    "b2_ofdmsync_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float * 4)]),
    "b2_mctx_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, C.c_uint, _vp, C.c_int, C.POINTER(_vp)]),
    "b2_mctx_destroy": (C.c_int, [_vp]),
    "b2_mctx_reset": (C.c_int, [_vp]),
    "b2_mctx_nco_advance": (C.c_int, [_vp, C.c_int64]),
    "b2_mctx_is_ready": (C.c_int, [_vp, C.c_uint, C.POINTER(C.c_int)]),
    "b2_mctx_update": (C.c_int, [_vp, C.c_uint, _vp, _vp, C.c_uint, C.c_int, C.c_int, C.c_int]),
    "b2_mctx_update_many": (C.c_int, [_vp, C.c_uint, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint)]),
    "b2_mctx_generate": (C.c_int, [_vp, _vp, _sz]),
    "b2_mctx_generate_device": (C.c_int, [_vp, _vp, _sz]),
    "b2_mctx_calls_to_boundary": (C.c_int, [_vp, C.POINTER(_sz)]),
    "b2_mctx_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float * 4)]),
    "b2_ofdmgen_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, _vp, C.c_int, C.POINTER(_vp)]),
    "b2_ofdmgen_destroy": (C.c_int, [_vp]),
    "b2_ofdmgen_reset": (C.c_int, [_vp]),
    "b2_ofdmgen_is_assembled": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "b2_ofdmgen_assemble": (C.c_int, [_vp, _vp, _vp, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint)]),
    "b2_ofdmgen_write": (C.c_int, [_vp, _vp, C.c_uint, C.POINTER(C.c_int)]),
    "b2_msresamp_create": (C.c_int, [C.c_float, C.c_float, C.c_int, C.POINTER(_vp)]),
    "b2_msresamp_destroy": (C.c_int, [_vp]),
    "b2_msresamp_reset": (C.c_int, [_vp]),
    "b2_msresamp_execute": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "b2_msresamp_execute_device": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "b2_msresamp_execute_to_device": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "b2_mcrx_set_resampler": (C.c_int, [_vp, C.c_float, C.c_float]),
    "b2_mcrx_shard_create": (C.c_int, [C.c_uint, C.c_uint, C.c_uint, C.c_uint, _vp, C.c_int, C.c_uint, C.c_uint, _sz, _sz, _vp, _vp,
                                       C.POINTER(_vp)]),
    "b2_mcrx_shard_destroy": (C.c_int, [_vp]),
    "b2_mcrx_shard_export": (C.c_int, [_vp, _vp]),
    "b2_mcrx_shard_connect": (C.c_int, [_vp, _vp, _sz]),
    "b2_mcrx_shard_begin": (C.c_int, [_vp]),
    "b2_mcrx_shard_stage1": (C.c_int, [_vp, _vp, C.c_uint64]),
    "b2_mcrx_shard_stage2": (C.c_int, [_vp, C.c_uint64]),
    "b2_mcrx_shard_end": (C.c_int, [_vp]),
    "b2_mcrx_shard_poll": (C.c_int, [_vp, _vp, _sz, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz)]),
    "b2_mcrx_shard_reset": (C.c_int, [_vp]),
    "b2_mcrx_shard_host_results": (C.c_int, [_vp, C.c_int]),
    "b2_memcpy_async": (C.c_int, [_vp, _vp, _sz, _vp]),
    "b2_mcrx_shard_poll_view": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz)]),
    "b2_mcrx_shard_pack_results": (C.c_int, [_vp, _vp, _sz, C.c_uint64, C.POINTER(_sz), C.POINTER(_sz)]),
}


def lib():
    """load libb200ofdm.so; fails loudly when the CUDA library has not been built"""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % path)
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in _PROTOS.items():
            f = getattr(L, name)            # AttributeError if the library lacks a declared symbol
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise B2Error(rc, lib().b2_last_error().decode("utf-8", "replace"))


def _ptr(x):
    """host numpy array or raw device pointer (int) -> void*"""
    if isinstance(x, (int, np.integer)):
        return C.c_void_p(int(x))
    return C.c_void_p(x.ctypes.data)


class _FrameSource:
    _prefix = None

    def _fn(self, name):
        return getattr(lib(), self._prefix + name)

    def poll(self):
        """-> (records as a FRAME_DTYPE array, payload bytes); clears the queue"""
        n, nb = _sz(0), _sz(0)
        _check(self._fn("poll")(self.h, None, 0, C.byref(n), None, 0, C.byref(nb)))
        recs = np.empty(n.value, FRAME_DTYPE)
        pl = np.empty(max(nb.value, 1), np.uint8)
        if n.value:
            _check(self._fn("poll")(self.h, recs.ctypes.data, n.value, C.byref(n), pl.ctypes.data, len(pl), C.byref(nb)))
        return recs, pl[:nb.value]

    def poll_view(self):
        """zero-copy poll: arrays alias the library's buffers and are valid only until the next
        execute / poll on this handle"""
        pr, pp, n, nb = _vp(), _vp(), _sz(0), _sz(0)
        _check(self._fn("poll_view")(self.h, C.byref(pr), C.byref(n), C.byref(pp), C.byref(nb)))
        if n.value == 0:
            return np.zeros(0, FRAME_DTYPE), np.zeros(0, np.uint8)
        recs = np.frombuffer((C.c_char * (n.value * FRAME_DTYPE.itemsize)).from_address(pr.value), dtype=FRAME_DTYPE)
        pl = np.frombuffer((C.c_char * max(nb.value, 1)).from_address(pp.value), dtype=np.uint8)[:nb.value] if nb.value else np.zeros(0, np.uint8)
        return recs, pl

    def last_timing(self):
        ms = (C.c_float * 4)()
        _check(self._fn("last_timing")(self.h, C.byref(ms)))
        return [float(v) for v in ms]

    def last_launches(self):
        """(kernels launched, pipeline chunks) of the last execute call (multichannelrx only)"""
        k, c = C.c_uint(0), C.c_uint(0)
        _check(lib().b2_mcrx_last_launches(self.h, C.byref(k), C.byref(c)))
        return int(k.value), int(c.value)

    def reset(self):
        _check(self._fn("reset")(self.h))

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultichannelRx(_FrameSource):
    """b2_mcrx_*: N-channel OFDM receiver (multichannelrx, lib/multichannelrx.cc)"""
    _prefix = "b2_mcrx_"

    def __init__(self, num_channels, M, cp_len, taper_len, p=None, device=0, max_batch=0):
        self.N, self.M, self.cp, self.taper = num_channels, M, cp_len, taper_len
        h = _vp()
        p_arr = None if p is None else np.ascontiguousarray(p, np.uint8)      # kept alive until the call returns
        pp = None if p_arr is None else p_arr.ctypes.data
        _check(lib().b2_mcrx_create(num_channels, M, cp_len, taper_len, pp, device, max_batch, C.byref(h)))
        self.h = h

    def execute(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        _check(lib().b2_mcrx_execute(self.h, x.ctypes.data, len(x)))

    def execute_device(self, dev_ptr, n):
        _check(lib().b2_mcrx_execute_device(self.h, _ptr(dev_ptr), n))

    def set_resampler(self, rate, As=60.0):
        """msresamp_crcf ahead of the NCO / channelizer (rate = 0 removes it)"""
        _check(lib().b2_mcrx_set_resampler(self.h, float(rate), float(As)))

    def tap_symbols(self, enable=True, max_symbols=1 << 16):
        _check(lib().b2_mcrx_tap_symbols(self.h, int(enable), max_symbols))

    def read_symbols(self):
        n = _sz(0)
        _check(lib().b2_mcrx_read_symbols(self.h, None, None, None, 0, C.byref(n)))
        ch = np.zeros(n.value, np.uint32)
        idx = np.zeros(n.value, np.uint64)
        X = np.zeros((n.value, self.M), np.complex64)
        if n.value:
            _check(lib().b2_mcrx_read_symbols(self.h, ch.ctypes.data, idx.ctypes.data, X.ctypes.data, n.value, C.byref(n)))
        return ch, idx, X

    def read_channelizer(self):
        nb = _sz(0)
        _check(lib().b2_mcrx_read_channelizer(self.h, None, 0, C.byref(nb)))
        out = np.zeros((self.N, nb.value), np.complex64)
        if nb.value:
            _check(lib().b2_mcrx_read_channelizer(self.h, out.ctypes.data, out.size, C.byref(nb)))
        return out

    def stream(self):
        return lib().b2_mcrx_stream(self.h)

    def channelize_device(self, x_ptr, n_blocks, sample_offset, out_ptr, out_stride):
        _check(lib().b2_mcrx_channelize_device(self.h, _ptr(x_ptr), n_blocks, sample_offset, _ptr(out_ptr), out_stride))

    def sync_device(self, in_ptr, n, in_stride):
        _check(lib().b2_mcrx_sync_device(self.h, _ptr(in_ptr), n, in_stride))


class OfdmSync(_FrameSource):
    """b2_ofdmsync_*: `streams` independent ofdmflexframesync instances (lib/ofdmtxrx.cc:91,625)"""
    _prefix = "b2_ofdmsync_"

    def __init__(self, M, cp_len, taper_len, p=None, streams=1, device=0, max_batch=0):
        self.M, self.cp, self.taper, self.streams = M, cp_len, taper_len, streams
        h = _vp()
        p_arr = None if p is None else np.ascontiguousarray(p, np.uint8)      # kept alive until the call returns
        pp = None if p_arr is None else p_arr.ctypes.data
        _check(lib().b2_ofdmsync_create(M, cp_len, taper_len, pp, streams, device, max_batch, C.byref(h)))
        self.h = h

    def execute(self, x):
        """x: [streams, n] complex64"""
        x = np.ascontiguousarray(x, np.complex64).reshape(self.streams, -1)
        _check(lib().b2_ofdmsync_execute(self.h, x.ctypes.data, x.shape[1]))

    def execute_device(self, dev_ptr, n, stride):
        _check(lib().b2_ofdmsync_execute_device(self.h, _ptr(dev_ptr), n, stride))


class _Handle:
    _destroy = None

    def close(self):
        if getattr(self, "h", None):
            getattr(lib(), self._destroy)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultichannelTx(_Handle):
    """b2_mctx_*: N-channel OFDM transmitter (multichanneltx, lib/multichanneltx.cc)"""
    _destroy = "b2_mctx_destroy"

    def __init__(self, num_channels, M, cp_len, taper_len, p=None, device=0):
        self.N, self.M, self.cp, self.taper = num_channels, M, cp_len, taper_len
        h = _vp()
        p_arr = None if p is None else np.ascontiguousarray(p, np.uint8)      # kept alive until the call returns
        pp = None if p_arr is None else p_arr.ctypes.data
        _check(lib().b2_mctx_create(num_channels, M, cp_len, taper_len, pp, device, C.byref(h)))
        self.h = h

    def reset(self):
        _check(lib().b2_mctx_reset(self.h))

    def is_ready(self, channel):
        r = C.c_int(0)
        _check(lib().b2_mctx_is_ready(self.h, channel, C.byref(r)))
        return r.value

    def update(self, channel, header, payload, mod, fec0, fec1):
        header = np.ascontiguousarray(header, np.uint8)
        payload = np.ascontiguousarray(payload, np.uint8)
        _check(lib().b2_mctx_update(self.h, channel, header.ctypes.data, payload.ctypes.data, len(payload), mod, fec0, fec1))

    def update_many(self, channels, headers, payloads, mod, fec0, fec1):
        """UpdateData for many channels in one call: headers [n, 8] uint8, payloads a list of uint8 arrays (or an [n, len]
        array); channels that are not ready are skipped.  Returns the number of channels taken."""
        channels = np.ascontiguousarray(channels, np.uint32)
        headers = np.ascontiguousarray(headers, np.uint8).reshape(len(channels), 8)
        if isinstance(payloads, np.ndarray) and payloads.ndim == 2:
            lens = np.full(len(channels), payloads.shape[1], np.uint32)
            flat = np.ascontiguousarray(payloads, np.uint8).reshape(-1)
        else:
            lens = np.array([len(p) for p in payloads], np.uint32)
            flat = np.concatenate([np.asarray(p, np.uint8) for p in payloads]) if len(payloads) else np.zeros(0, np.uint8)
        n = C.c_uint(0)
        _check(lib().b2_mctx_update_many(self.h, len(channels), channels.ctypes.data, headers.ctypes.data,
                                         flat.ctypes.data if len(flat) else None, lens.ctypes.data, mod, fec0, fec1, C.byref(n)))
        return n.value

    def generate(self, n_calls):
        out = np.zeros(n_calls * 2 * self.N, np.complex64)
        _check(lib().b2_mctx_generate(self.h, out.ctypes.data, n_calls))
        return out

    def generate_device(self, dev_ptr, n_calls):
        _check(lib().b2_mctx_generate_device(self.h, _ptr(dev_ptr), n_calls))

    def calls_to_boundary(self):
        n = _sz(0)
        _check(lib().b2_mctx_calls_to_boundary(self.h, C.byref(n)))
        return n.value

    def last_timing(self):
        ms = (C.c_float * 4)()
        _check(lib().b2_mctx_last_timing(self.h, C.byref(ms)))
        return [float(v) for v in ms]


class OfdmGen(_Handle):
    """b2_ofdmgen_*: one ofdmflexframegen (lib/ofdmtxrx.cc:79-84,314-328)"""
    _destroy = "b2_ofdmgen_destroy"

    def __init__(self, M, cp_len, taper_len, p=None, device=0):
        self.M, self.cp, self.taper = M, cp_len, taper_len
        h = _vp()
        p_arr = None if p is None else np.ascontiguousarray(p, np.uint8)      # kept alive until the call returns
        pp = None if p_arr is None else p_arr.ctypes.data
        _check(lib().b2_ofdmgen_create(M, cp_len, taper_len, pp, device, C.byref(h)))
        self.h = h

    def is_assembled(self):
        r = C.c_int(0)
        _check(lib().b2_ofdmgen_is_assembled(self.h, C.byref(r)))
        return r.value

    def assemble(self, header, payload, check, fec0, fec1, mod):
        header = np.ascontiguousarray(header, np.uint8)
        payload = np.ascontiguousarray(payload, np.uint8)
        n = C.c_uint(0)
        _check(lib().b2_ofdmgen_assemble(self.h, header.ctypes.data, payload.ctypes.data, len(payload), check, fec0, fec1, mod, C.byref(n)))
        return n.value

    def write(self, n_symbols):
        out = np.zeros(n_symbols * (self.M + self.cp), np.complex64)
        last = C.c_int(0)
        _check(lib().b2_ofdmgen_write(self.h, out.ctypes.data, n_symbols, C.byref(last)))
        return out, last.value


class MsResamp(_Handle):
    """b2_msresamp_*: msresamp_crcf (src/flexframe_rx.cc:179,240)"""
    _destroy = "b2_msresamp_destroy"

    def __init__(self, rate, As=60.0, device=0):
        self.rate = rate
        h = _vp()
        _check(lib().b2_msresamp_create(rate, As, device, C.byref(h)))
        self.h = h

    def reset(self):
        _check(lib().b2_msresamp_reset(self.h))

    def execute(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        y = np.zeros(int(len(x) * float(self.rate) * 1.01) + 64, np.complex64)
        ny = _sz(0)
        _check(lib().b2_msresamp_execute(self.h, x.ctypes.data, len(x), y.ctypes.data, len(y), C.byref(ny)))
        return y[:ny.value]

    def execute_device(self, x_ptr, nx, y_ptr, y_cap):
        """device pointers in and out; returns the number of output samples"""
        ny = _sz(0)
        _check(lib().b2_msresamp_execute_device(self.h, _vp(x_ptr), nx, _vp(y_ptr), y_cap, C.byref(ny)))
        return ny.value

    def execute_to_device(self, x, y_ptr, y_cap):
        """host samples in, device pointer out"""
        x = np.ascontiguousarray(x, np.complex64)
        ny = _sz(0)
        _check(lib().b2_msresamp_execute_to_device(self.h, x.ctypes.data, len(x), _vp(y_ptr), y_cap, C.byref(ny)))
        return ny.value
