"""On-disk sample format of the offline UHD stand-in (include/uhd/usrp/multi_usrp.hpp): interleaved
little-endian complex float32 -- the layout of the uhd::io_type_t::COMPLEX_FLOAT32 buffers the
reference hands to send()/recv() (lib/ofdmtxrx.cc:338, src/multichannel_rx.cc:200-211) -- plus a
SigMF-style sidecar "<file>.sigmf-meta" carrying what is needed to replay a capture."""
import json
import os

import numpy as np

DATATYPE = "cf32_le"


def meta_path(path):
    return str(path) + ".sigmf-meta"


def write_capture(path, x, sample_rate, frequency=0.0, description="", **extra):
    """samples -> <path>, settings -> <path>.sigmf-meta; extra keys are stored as "b2:<key>" """
    x = np.ascontiguousarray(x, dtype="<c8")
    x.tofile(str(path))
    g = {"core:datatype": DATATYPE, "core:version": "1.0.0", "core:sample_rate": float(sample_rate),
         "core:description": description, "b2:samples": int(len(x))}
    for k, v in extra.items():
        g["b2:" + k] = v
    meta = {"global": g, "captures": [{"core:sample_start": 0, "core:frequency": float(frequency)}], "annotations": []}
    with open(meta_path(path), "w") as f:
        json.dump(meta, f, indent=2)
    return meta


def read_meta(path):
    """the sidecar of a capture, or None when it has none"""
    p = meta_path(path)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        meta = json.load(f)
    dt = meta.get("global", {}).get("core:datatype", DATATYPE)
    if dt != DATATYPE:
        raise ValueError("capture %s holds %s samples; only %s is supported" % (path, dt, DATATYPE))
    return meta


def read_capture(path, count=-1, offset=0):
    """(samples as complex64, sidecar or None); offset / count in samples"""
    meta = read_meta(path)
    x = np.fromfile(str(path), dtype="<c8", count=count, offset=8 * offset)
    return x.astype(np.complex64, copy=False), meta


def resample_ratio(meta, wanted_rate):
    """rate to give b2_mcrx_set_resampler / msresamp so that a capture plays at wanted_rate"""
    return float(wanted_rate) / float(meta["global"]["core:sample_rate"])
