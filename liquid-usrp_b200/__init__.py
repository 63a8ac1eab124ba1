"""liquid-usrp_b200 -- B200 (sm_100a) implementation of liquid-usrp's multichannel OFDM DSP path.

The product is native: CUDA kernels + a C ABI (libb200ofdm.so, include/b200_ofdm.h) and the
reference-compatible C++ classes on top (libliquidusrp_b200.so).  This Python package is a thin
ctypes view of the C ABI used by tests/ and bench.py; there is no Python or CPU fallback -- if the
shared library is missing, importing `capi` raises.

The directory name carries a hyphen (it mirrors the reference's name), so import it with
    importlib.import_module("liquid-usrp_b200")
"""
from .capi import (B2Error, FRAME_DTYPE, MsResamp, MultichannelRx, MultichannelTx, OfdmGen, OfdmSync,  # noqa: F401
                   lib, lib_path)
from . import capture  # noqa: F401
