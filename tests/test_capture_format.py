"""On-disk sample format: cf32 interleaved little-endian files plus a SigMF-style
sidecar, written by the offline UHD stand-in (include/uhd/usrp/multi_usrp.hpp) and by
liquid-usrp_b200/capture.py, each readable by the other.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

PROG = r'''
#include <uhd/usrp/multi_usrp.hpp>
#include <complex>
#include <vector>
int main(int argc, char ** argv)
{
    uhd::device_addr_t addr;
    uhd::usrp::multi_usrp::sptr usrp = uhd::usrp::multi_usrp::make(addr);
    if (argc > 1 && argv[1][0] == 't') {                 // transmit 1000 samples the way lib/ofdmtxrx.cc:338 does
        usrp->set_tx_rate(2.5e6); usrp->set_tx_freq(462e6); usrp->set_tx_gain(-3.0);
        std::vector<std::complex<float> > buf(250);
        uhd::tx_metadata_t md;
        for (int b = 0; b < 4; b++) {
            for (int i = 0; i < 250; i++) buf[i] = std::complex<float>(b * 250 + i, -(b * 250 + i));
            usrp->get_device()->send(&buf.front(), buf.size(), md, uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
        }
    } else {                                             // receive: print the sum of what the file held
        usrp->set_rx_rate(2.0e6);
        usrp->issue_stream_cmd(uhd::stream_cmd_t::STREAM_MODE_START_CONTINUOUS);
        std::vector<std::complex<float> > buf(300);
        uhd::rx_metadata_t md;
        double acc = 0; size_t n, total = 0;
        while ((n = usrp->get_device()->recv(&buf.front(), buf.size(), md, uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::RECV_MODE_FULL_BUFF)) > 0) {
            for (size_t i = 0; i < n; i++) acc += buf[i].real();
            total += n;
        }
        printf("received %zu samples, sum %.1f\n", total, acc);
    }
    return 0;
}
'''


def _build(tmp_path):
    src = tmp_path / "stub_io.cc"
    src.write_text(PROG)
    exe = tmp_path / "stub_io"
    subprocess.check_call(["g++", "-std=gnu++17", "-O1", "-w", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    return str(exe)


def test_stub_writes_cf32_and_sidecar_that_python_reads(tmp_path):
    from b2 import pkg
    exe = _build(tmp_path)
    f = tmp_path / "air.cf32"
    env = dict(os.environ, B2_UHD_TX_FILE=str(f), B2_UHD_NOTE="N=4 M=64 cp=16 taper=4")
    subprocess.check_call([exe, "t"], env=env)
    x, meta = pkg.capture.read_capture(f)
    assert len(x) == 1000 and x.dtype == np.complex64
    assert np.array_equal(x.real, np.arange(1000, dtype=np.float32)) and np.array_equal(x.imag, -np.arange(1000, dtype=np.float32))
    g = meta["global"]
    assert g["core:datatype"] == "cf32_le" and g["core:sample_rate"] == 2.5e6 and g["b2:samples"] == 1000
    assert g["core:description"] == "N=4 M=64 cp=16 taper=4" and g["b2:tx_gain_db"] == -3.0
    assert meta["captures"][0]["core:frequency"] == 462e6
    assert abs(pkg.capture.resample_ratio(meta, 2.0e6) - 0.8) < 1e-12
    # the receive side of the stand-in reads the same file and notices the rate mismatch through the sidecar
    r = subprocess.run([exe, "r"], env=dict(os.environ, B2_UHD_RX_FILE=str(f)), capture_output=True, text=True)
    assert "received 1000 samples, sum 499500.0" in r.stdout, r.stdout
    assert "capture was taken at 2.5e+06 S/s" in r.stderr and "resample by 0.8" in r.stderr, r.stderr


def test_python_capture_is_read_by_the_stub(tmp_path):
    from b2 import pkg
    exe = _build(tmp_path)
    f = tmp_path / "synth.cf32"
    x = (np.arange(777) + 1j * np.ones(777)).astype(np.complex64)
    meta = pkg.capture.write_capture(f, x, 2.0e6, frequency=915e6, description="synthetic", num_channels=8, M=64)
    assert os.path.getsize(f) == 777 * 8
    assert json.load(open(pkg.capture.meta_path(f))) == meta and meta["global"]["b2:M"] == 64
    y, m2 = pkg.capture.read_capture(f, count=100, offset=10)
    assert np.array_equal(y, x[10:110]) and m2 == meta
    r = subprocess.run([exe, "r"], env=dict(os.environ, B2_UHD_RX_FILE=str(f)), capture_output=True, text=True)
    assert "received 777 samples, sum %.1f" % float(np.arange(777).sum()) in r.stdout, r.stdout
    assert "capture was taken" not in r.stderr            # same rate: no warning
    # a file without a sidecar is still a valid capture
    g = tmp_path / "bare.cf32"
    x.tofile(str(g))
    z, m3 = pkg.capture.read_capture(g)
    assert m3 is None and np.array_equal(z, x)
