"""CPU-side checks of the drop-in boundary: the C ABI library loads and exports every symbol
include/b200_ofdm.h declares, the reference's own programs compile and link UNMODIFIED against
this repo's headers and libraries, constructor errors behave like the reference's (`throw 0`),
and with no CUDA device the product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "liquid-usrp_b200")
REF = "/root/reference"


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


SHIM = os.path.join(ROOT, "tests", "shim", "libmcshim_b200.so")      # test driver of the host classes (not in the product library)


@pytest.fixture(scope="module")
def built():
    if not (os.path.exists(os.path.join(PKG, "libb200ofdm.so")) and os.path.exists(os.path.join(PKG, "libliquidusrp_b200.so")) and os.path.exists(SHIM)):
        subprocess.check_call(["make", "-C", PKG, "-j8"])
    return PKG


def test_c_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "b200_ofdm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    lib = C.CDLL(os.path.join(built, "libb200ofdm.so"))
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    lib.b2_version.restype = C.c_char_p
    assert b"sm_100a" in lib.b2_version()


def test_python_binding_covers_the_header(built):
    from b2 import pkg
    pkg.lib()                       # raises AttributeError on any prototype without a symbol
    hdr = open(os.path.join(ROOT, "include", "b200_ofdm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    import importlib
    capi = importlib.import_module("liquid-usrp_b200.capi")
    assert declared <= set(capi._PROTOS), sorted(declared - set(capi._PROTOS))


def test_host_library_exports_reference_class_symbols(built):
    out = subprocess.check_output(["nm", "-DC", os.path.join(built, "libliquidusrp_b200.so")], text=True)
    for sym in ("multichannelrx::multichannelrx(unsigned int, unsigned int, unsigned int, unsigned int, unsigned char*, void**, int (**)(",
                "multichannelrx::Execute(std::complex<float>*, unsigned int)", "multichannelrx::Reset()",
                "multichanneltx::multichanneltx(unsigned int, unsigned int, unsigned int, unsigned int, unsigned char*)",
                "multichanneltx::IsChannelReadyForData(unsigned int)", "multichanneltx::GenerateSamples(std::complex<float>*)",
                "multichanneltx::UpdateData(unsigned int, unsigned char*, unsigned char*, unsigned int, int, int, int)",
                "ofdmtxrx::transmit_packet(unsigned char*, unsigned char*, unsigned int, int, int, int)",
                "ofdmtxrx::write_symbol()", "ofdmtxrx::start_rx()", "ofdmtxrx_rx_worker(void*)", "ofdmtxrx_rx_worker_blocking(void*)",
                "timer_create()", "ofdmflexframegen_write", "ofdmflexframesync_execute", "liquid_getopt_str2mod", "msresamp_crcf_execute"):
        assert sym in out, sym


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("prog", ["multichannel_rx", "multichannel_tx", "ofdmflexframe_rx", "ofdmflexframe_tx",
                                  "fullduplex_txrx", "halfduplex_txrx", "multichannel_txrx"])
def test_reference_programs_link_unmodified(built, prog, tmp_path):
    out = tmp_path / prog
    cmd = ["g++", "-std=gnu++17", "-O2", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "host"),
           os.path.join(REF, "src", prog + ".cc"), "-o", str(out), "-L" + PKG, "-lliquidusrp_b200", "-lb200ofdm",
           "-Wl,-rpath," + PKG, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    # usage text comes from the reference's own main(); -h needs no radio and no GPU
    r = subprocess.run([str(out), "-h"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and ("usage" in (r.stdout + r.stderr).lower() or prog in r.stdout), (r.returncode, r.stdout[-300:], r.stderr[-300:])


def test_constructor_errors_throw_like_the_reference(built):
    from refmc import McLib, McRx, McTx
    L = McLib(SHIM)
    for args in ((0, 64, 16, 4), (2, 6, 2, 0), (2, 64, 0, 0), (2, 64, 4, 8)):
        with pytest.raises(ValueError):
            McRx(L, *args)
        with pytest.raises(ValueError):
            McTx(L, *args)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a CUDA device")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback(built):
    from b2 import pkg
    from refmc import McLib, McRx
    with pytest.raises(pkg.B2Error) as e:
        pkg.MultichannelRx(8, 64, 16, 4)
    assert e.value.code == -3 and "CUDA" in str(e.value)
    with pytest.raises(pkg.B2Error):
        pkg.MultichannelTx(8, 64, 16, 4)
    with pytest.raises(pkg.B2Error):
        pkg.MsResamp(1.07)
    L = McLib(SHIM)
    with pytest.raises(ValueError):         # the class throws 0 when the device library cannot start
        McRx(L, 8, 64, 16, 4)


def test_product_never_touches_the_oracle(built):
    """the oracle is test infrastructure: nothing under liquid-usrp_b200/ may reference it"""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cc", ".py")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"liborc|oracle/|orc_[a-z]+\(|libref_mc", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    for so in ("libb200ofdm.so", "libliquidusrp_b200.so"):
        deps = subprocess.check_output(["ldd", os.path.join(built, so)], text=True)
        assert "liborc" not in deps and "libref_mc" not in deps
